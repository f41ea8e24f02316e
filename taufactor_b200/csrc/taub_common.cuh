// taub_common.cuh -- shared host/device helpers for libtaub200 (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "taub200.h"

#if defined(__CUDA_ARCH__) && (__CUDA_ARCH__ < 1000)
#error "libtaub200 is written for sm_100a (B200) only"
#endif

namespace taub {

constexpr int G = TAUB_GHOST;
constexpr int COL0 = TAUB_COL0;

void set_error(const char *fmt, ...);
void count_launch(int n = 1);   // bumps the counter behind taub_launch_count()

#define TAUB_CUDA(call)                                                                    \
    do {                                                                                   \
        cudaError_t e_ = (call);                                                           \
        if (e_ != cudaSuccess) {                                                           \
            ::taub::set_error("%s:%d: %s -> %s", __FILE__, __LINE__, #call,                \
                              cudaGetErrorString(e_));                                     \
            return TAUB_ERR_CUDA;                                                          \
        }                                                                                  \
    } while (0)

#define TAUB_REQUIRE(cond, ...)                                                            \
    do {                                                                                   \
        if (!(cond)) {                                                                     \
            ::taub::set_error(__VA_ARGS__);                                                \
            return TAUB_ERR_ARG;                                                           \
        }                                                                                  \
    } while (0)

static inline int ceil_div(int a, int b) { return (a + b - 1) / b; }
static inline int64_t ceil_div64(int64_t a, int64_t b) { return (a + b - 1) / b; }

// Interior float4 groups per row: interior columns [4, 4+Nz) -> groups [1, 1+ngroups).
static inline __host__ __device__ int interior_groups(int Nz) { return (Nz + 3) >> 2; }

// ------------------------------------------------------------------------------------------
// Exact division by a small neighbour count.
//
// The reference divides the neighbour sum by `factor` in {1..8, inf} with an IEEE fp32 divide
// (taufactor.py:177).  For a divisor c with correctly rounded reciprocal r = RN(1/c) the
// Markstein sequence  q0 = RN(s*r); rem = s - q0*c (exact, one FMA); q = RN(q0 + rem*r)
// returns the correctly rounded quotient for every normal s (checked exhaustively over all 2^23
// mantissas for c = 1..8).  Tiny non-zero |s| (< 2^-100, where the remainder could go subnormal)
// takes the __fdiv_rn path, so the result is bit-identical to IEEE division for every finite s.
// Table entry for code 0 ("factor = inf": non-conductive voxel or no conductive neighbour) is
// (c, r) = (0, 0), which yields q = 0 = s / inf without a special case.
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ float2 div_entry(int code)
{
    float c = (code >= 1 && code <= 8) ? (float)code : 0.0f;
    float r = (c > 0.0f) ? __frcp_rn(c) : 0.0f;
    return make_float2(c, r);
}

__device__ __forceinline__ float div_small(float s, float2 cr)
{
    float q0 = __fmul_rn(s, cr.y);
    float rem = __fmaf_rn(-q0, cr.x, s);
    float q = __fmaf_rn(rem, cr.y, q0);
    // rare path: 0 < |s| < 2^-100 (one shift-add and one unsigned compare: 2*bits drops the sign,
    // minus 1 wraps +-0 to the top).  s == 0 -- every voxel inside the solid phase -- stays on the
    // fast path, which returns the exact 0; inf / NaN propagate as NaN (the field has diverged).
    const unsigned u = __float_as_uint(s) * 2u - 1u;
    if (u < (27u << 24) - 1u) q = (cr.x > 0.0f) ? __fdiv_rn(s, cr.x) : 0.0f;
    return q;
}

// One binary-solver voxel update, the reference's op order (taufactor.py:97-102, :177-181):
// s = ((((x+ + x-) + y+) + y-) + z+) + z-;  f += omega * (s / nn - f).  No FMA contraction.
__device__ __forceinline__ float sor_binary(float c, float xp, float xm, float yp, float ym,
                                            float zp, float zm, float2 cr, float omega)
{
    float s = __fadd_rn(xp, xm);
    s = __fadd_rn(s, yp);
    s = __fadd_rn(s, ym);
    s = __fadd_rn(s, zp);
    s = __fadd_rn(s, zm);
    float d = __fsub_rn(div_small(s, cr), c);
    d = __fmul_rn(d, omega);
    return __fadd_rn(c, d);
}

// Colour updates of two voxels of a float4 group.  "xz" rows update components x and z (their z
// neighbours are y, w and the scalar zs = .w of the group on the left); "yw" rows update y and w
// (zs = .x of the group on the right).  xp/xm: x neighbours, up/dn: y+1 / y-1 neighbours.
__device__ __forceinline__ void update_xz(float4 &c, const float4 &xp, const float4 &xm, const float4 &up,
                                          const float4 &dn, float zs, unsigned code, const float2 *s_div, float omega)
{
    c.x = sor_binary(c.x, xp.x, xm.x, up.x, dn.x, c.y, zs, s_div[code & 15u], omega);
    c.z = sor_binary(c.z, xp.z, xm.z, up.z, dn.z, c.w, c.y, s_div[(code >> 8) & 15u], omega);
}
__device__ __forceinline__ void update_yw(float4 &c, const float4 &xp, const float4 &xm, const float4 &up,
                                          const float4 &dn, float zs, unsigned code, const float2 *s_div, float omega)
{
    c.y = sor_binary(c.y, xp.y, xm.y, up.y, dn.y, c.z, c.x, s_div[(code >> 4) & 15u], omega);
    c.w = sor_binary(c.w, xp.w, xm.w, up.w, dn.w, zs, c.z, s_div[(code >> 12) & 15u], omega);
}

// One multi-phase voxel update (taufactor.py:606-613, :598-603): each neighbour times its face
// conductance (separately rounded), summed left to right; prefactor = sum of the six face
// conductances in the reference's order (+ the Dirichlet face once more on the first / last
// plane), 0 -> inf; IEEE division.
__device__ __forceinline__ float sor_multi(float c, float xp, float xm, float yp, float ym,
                                           float zp, float zm, float wxp, float wxm, float wyp,
                                           float wym, float wzp, float wzm, bool first, bool last,
                                           float omega)
{
    float s = __fadd_rn(__fmul_rn(xp, wxp), __fmul_rn(xm, wxm));
    s = __fadd_rn(s, __fmul_rn(yp, wyp));
    s = __fadd_rn(s, __fmul_rn(ym, wym));
    s = __fadd_rn(s, __fmul_rn(zp, wzp));
    s = __fadd_rn(s, __fmul_rn(zm, wzm));
    float fac = __fadd_rn(wxm, wxp);
    fac = __fadd_rn(fac, wym);
    fac = __fadd_rn(fac, wyp);
    fac = __fadd_rn(fac, wzm);
    fac = __fadd_rn(fac, wzp);
    if (first) fac = __fadd_rn(fac, wxm);
    if (last) fac = __fadd_rn(fac, wxp);
    if (fac == 0.0f) fac = __int_as_float(0x7f800000);
    float d = __fsub_rn(__fdiv_rn(s, fac), c);
    d = __fmul_rn(d, omega);
    return __fadd_rn(c, d);
}

__device__ __forceinline__ int wrap(int a, int n)
{
    a %= n;
    return a < 0 ? a + n : a;
}

}  // namespace taub
