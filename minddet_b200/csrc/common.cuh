// common.cuh -- device helpers shared by the region-path kernels (sm_100a only).
//
// Everything that feeds an integer decision (top-k order, NMS keep, assignment, RoI level) is written
// with explicitly rounded intrinsics (__fmul_rn/__fadd_rn/__fdiv_rn never contract into FMA) so that
// the result is bit-identical to a strict-fp32 CPU evaluation of the same expression tree
// (SURVEY.md section 7.2 "Bit-exact NMS under fp32").  The library is also compiled with -fmad=false.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#define MD_DEVINL __device__ __forceinline__

namespace md {

constexpr int kMaxLevels = 8;

// ---- shared memory through explicit 32-bit shared-space addresses ------------------------------------------
// In a cluster kernel nvcc rebuilds the shared-window base from SR_CgaCtaId (S2R / S2UR) in front of nearly every
// access made through a pointer, also inside hot loops.  smem_addr() computes the address once and hides it behind
// an opaque asm so it stays in a register; the helpers below take such addresses.
MD_DEVINL uint32_t smem_addr(const void *p)
{
    uint32_t a = (uint32_t)__cvta_generic_to_shared(p);
    asm volatile("" : "+r"(a));
    return a;
}
MD_DEVINL uint32_t lds32(uint32_t a) { uint32_t v; asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(a)); return v; }
MD_DEVINL uint4 lds128(uint32_t a)
{
    uint4 v;
    asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(a));
    return v;
}
MD_DEVINL unsigned long long lds64(uint32_t a) { unsigned long long v; asm volatile("ld.shared.u64 %0, [%1];" : "=l"(v) : "r"(a)); return v; }
MD_DEVINL void sts64(uint32_t a, unsigned long long v) { asm volatile("st.shared.b64 [%0], %1;" :: "r"(a), "l"(v) : "memory"); }
MD_DEVINL void sts128(uint32_t a, uint4 v)
{
    asm volatile("st.shared.v4.u32 [%0], {%1, %2, %3, %4};" :: "r"(a), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}
MD_DEVINL uint32_t atoms_add(uint32_t a, uint32_t v)
{
    uint32_t r;
    asm volatile("atom.shared.add.u32 %0, [%1], %2;" : "=r"(r) : "r"(a), "r"(v) : "memory");
    return r;
}
MD_DEVINL void reds_add(uint32_t a, uint32_t v) { asm volatile("red.shared.add.u32 [%0], %1;" :: "r"(a), "r"(v) : "memory"); }

// ---- explicitly rounded arithmetic ------------------------------------------------------------
MD_DEVINL float mul(float a, float b) { return __fmul_rn(a, b); }
MD_DEVINL float add(float a, float b) { return __fadd_rn(a, b); }
MD_DEVINL float sub(float a, float b) { return __fsub_rn(a, b); }
MD_DEVINL float div(float a, float b) { return __fdiv_rn(a, b); }

// Deterministic fp32 exp (oracle: o_exp): round-to-nearest range reduction through the 1.5*2^23 magic constant,
// Cody-Waite with two FMAs, degree-5 polynomial in Horner form with FMAs (an fma.rn is ONE rounding of the exact
// product-sum, so it is reproduced bit for bit by fmaf() on the CPU), exact 2^k scaling.  <= 2 ulp.
// Input clamped to [-87, 88].  18 instructions, all on the FMA/ALU pipes (no F2I / FRND).
MD_DEVINL float exact_exp(float x)
{
    x = fminf(x, 88.0f);
    x = fmaxf(x, -87.0f);
    const float t = mul(x, 1.44269504088896341f);
    const float z = add(t, 12582912.0f);
    const float kf = sub(z, 12582912.0f);
    float r = __fmaf_rn(kf, -0.693359375f, x);
    r = __fmaf_rn(kf, 2.12194440e-4f, r);
    float p = 1.9875691500E-4f;
    p = __fmaf_rn(p, r, 1.3981999507E-3f);
    p = __fmaf_rn(p, r, 8.3334519073E-3f);
    p = __fmaf_rn(p, r, 4.1665795894E-2f);
    p = __fmaf_rn(p, r, 1.6666665459E-1f);
    p = __fmaf_rn(p, r, 5.0000001201E-1f);
    p = __fmaf_rn(p, mul(r, r), r);
    p = add(p, 1.0f);
    const int k = __float_as_int(z) - 0x4B400000;
    return mul(p, __int_as_float((k + 127) << 23));
}

// 1 / y through __frcp_rn: the correctly rounded reciprocal IS the correctly rounded quotient 1.0f / y of the oracle,
// at about half the instructions of the general __fdiv_rn sequence (the top-k evaluates 2.1 M sigmoids per step)
MD_DEVINL float exact_sigmoid(float x) { return __frcp_rn(add(1.0f, exact_exp(-x))); }

// monotone uint32 image of an fp32 value: larger float <=> larger key; -0.0 < +0.0
MD_DEVINL uint32_t score_key(float f)
{
    uint32_t b = __float_as_uint(f);
    return (b & 0x80000000u) ? ~b : (b | 0x80000000u);
}
MD_DEVINL float key_score(uint32_t k)
{
    uint32_t b = (k & 0x80000000u) ? (k & 0x7FFFFFFFu) : ~k;
    return __uint_as_float(b);
}

// ---- Philox-4x32-10, counter (n, stream, image, step), key = 64-bit seed; returns word 0 -----------
MD_DEVINL uint32_t philox_key(uint32_t n, uint32_t stream, uint32_t image, uint32_t seed_lo, uint32_t seed_hi, uint32_t step = 0u)
{
    uint32_t c0 = n, c1 = stream, c2 = image, c3 = step;
    uint32_t k0 = seed_lo, k1 = seed_hi;
#pragma unroll
    for (int r = 0; r < 10; r++) {
        uint32_t hi0 = __umulhi(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
        uint32_t hi1 = __umulhi(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
        uint32_t n0 = hi1 ^ c1 ^ k0, n2 = hi0 ^ c3 ^ k1;
        c0 = n0; c1 = lo1; c2 = n2; c3 = lo0;
        k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
    return c0;
}

// ---- box arithmetic ----------------------------------------------------------------------------
struct DecodeCfg {
    float img_h, img_w, mean[4], stdv[4], max_ratio;
};
MD_DEVINL DecodeCfg load_decode_cfg(const float *__restrict__ cfg)
{
    DecodeCfg c;
    c.img_h = __ldg(cfg + 0); c.img_w = __ldg(cfg + 1);
#pragma unroll
    for (int i = 0; i < 4; i++) { c.mean[i] = __ldg(cfg + 2 + i); c.stdv[i] = __ldg(cfg + 6 + i); }
    c.max_ratio = __ldg(cfg + 10);
    return c;
}

// legacy (+1) delta2bbox + clip; op order identical to oracle/region_oracle.c:decode_one
MD_DEVINL float4 decode_box(float4 a, float4 d, const DecodeCfg &c)
{
    float dx = add(mul(d.x, c.stdv[0]), c.mean[0]);
    float dy = add(mul(d.y, c.stdv[1]), c.mean[1]);
    float dw = add(mul(d.z, c.stdv[2]), c.mean[2]);
    float dh = add(mul(d.w, c.stdv[3]), c.mean[3]);
    dw = fminf(fmaxf(dw, -c.max_ratio), c.max_ratio);
    dh = fminf(fmaxf(dh, -c.max_ratio), c.max_ratio);
    float pw = add(sub(a.z, a.x), 1.0f);
    float ph = add(sub(a.w, a.y), 1.0f);
    float px = mul(add(a.x, a.z), 0.5f);
    float py = mul(add(a.y, a.w), 0.5f);
    float gw = mul(pw, exact_exp(dw));
    float gh = mul(ph, exact_exp(dh));
    float gx = add(px, mul(pw, dx));
    float gy = add(py, mul(ph, dy));
    float hw = mul(gw, 0.5f), hh = mul(gh, 0.5f);
    float x1 = add(sub(gx, hw), 0.5f);
    float y1 = add(sub(gy, hh), 0.5f);
    float x2 = sub(add(gx, hw), 0.5f);
    float y2 = sub(add(gy, hh), 0.5f);
    float mw = sub(c.img_w, 1.0f), mh = sub(c.img_h, 1.0f);
    float4 o;
    o.x = fminf(fmaxf(x1, 0.0f), mw);
    o.y = fminf(fmaxf(y1, 0.0f), mh);
    o.z = fminf(fmaxf(x2, 0.0f), mw);
    o.w = fminf(fmaxf(y2, 0.0f), mh);
    return o;
}

// legacy (+1) bbox2delta (FP tolerance: logf)
MD_DEVINL float4 encode_box(float4 p, float4 g, const float *mean, const float *stdv)
{
    float px = mul(add(p.x, p.z), 0.5f), py = mul(add(p.y, p.w), 0.5f);
    float pw = add(sub(p.z, p.x), 1.0f), ph = add(sub(p.w, p.y), 1.0f);
    float gx = mul(add(g.x, g.z), 0.5f), gy = mul(add(g.y, g.w), 0.5f);
    float gw = add(sub(g.z, g.x), 1.0f), gh = add(sub(g.w, g.y), 1.0f);
    float dx = div(sub(gx, px), pw), dy = div(sub(gy, py), ph);
    float dw = logf(div(gw, pw)), dh = logf(div(gh, ph));
    float4 o;
    o.x = div(sub(dx, mean[0]), stdv[0]);
    o.y = div(sub(dy, mean[1]), stdv[1]);
    o.z = div(sub(dw, mean[2]), stdv[2]);
    o.w = div(sub(dh, mean[3]), stdv[3]);
    return o;
}

// IoU with the legacy +off convention, zero unless iw>0 && ih>0 (oracle: iou_pair)
MD_DEVINL float iou_legacy(float4 a, float4 g, float garea, float off)
{
    float iw = add(sub(fminf(a.z, g.z), fmaxf(a.x, g.x)), off);
    if (!(iw > 0.0f)) return 0.0f;
    float ih = add(sub(fminf(a.w, g.w), fmaxf(a.y, g.y)), off);
    if (!(ih > 0.0f)) return 0.0f;
    float aarea = mul(add(sub(a.z, a.x), off), add(sub(a.w, a.y), off));
    float inter = mul(iw, ih);
    float ua = sub(add(aarea, garea), inter);
    return div(inter, ua);
}
MD_DEVINL float area_legacy(float4 g, float off)
{
    return mul(add(sub(g.z, g.x), off), add(sub(g.w, g.y), off));
}

// ---- 128-bit streaming loads / stores -----------------------------------------------------------
MD_DEVINL float4 ldg_stream(const float4 *p)
{
    float4 v;
    asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];"
                 : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p));
    return v;
}
MD_DEVINL void stg_stream(float4 *p, float4 v)
{
    asm volatile("st.global.L1::no_allocate.v4.f32 [%0], {%1,%2,%3,%4};"
                 :: "l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}

}  // namespace md
