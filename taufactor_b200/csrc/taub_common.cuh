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
extern thread_local bool g_fused_pdl;   // taub_fused.cu: launch the passes of taub_iterate as programmatic dependents
extern thread_local unsigned g_launch_cluster_x;   // taub_fused.cu: cluster width of the next launch_maybe_pdl (0 / 1: none)

// Launch with or without cudaLaunchAttributeProgrammaticStreamSerialization (g_fused_pdl).  A kernel launched
// through this must execute pdl_wait() in EVERY thread before its first access to global memory that an
// earlier grid of the stream may have written (and before any exit), and may call pdl_trigger() first.
template <typename... KArgs, typename... Args>
inline cudaError_t launch_maybe_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t s,
                                    Args &&...args)
{
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid;
    cfg.blockDim = block;
    cfg.dynamicSmemBytes = smem;
    cfg.stream = s;
    cudaLaunchAttribute attr[2];
    int na = 0;
    if (g_fused_pdl) {
        attr[na].id = cudaLaunchAttributeProgrammaticStreamSerialization;
        attr[na].val.programmaticStreamSerializationAllowed = 1;
        ++na;
    }
    if (g_launch_cluster_x > 1 && grid.x % g_launch_cluster_x == 0) {   // co-scheduled CTAs (no cluster barrier anywhere)
        attr[na].id = cudaLaunchAttributeClusterDimension;
        attr[na].val.clusterDim.x = g_launch_cluster_x;
        attr[na].val.clusterDim.y = 1;
        attr[na].val.clusterDim.z = 1;
        ++na;
    }
    cfg.attrs = attr;
    cfg.numAttrs = na;
    return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}
#ifdef __CUDACC__
// The next grid of the stream may be scheduled once every CTA of this one has got here (or exited).
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
// Block until the previous grid of the stream has completed and its writes are visible; returns at once in a
// grid that was not launched as a programmatic dependent.
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
#endif

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
// (c, r) = (0, 0), which yields q = 0 = s / inf without a special case.  The guard is evaluated once
// per group of updates (a branch per update would stop the compiler from interleaving the chains).
// ------------------------------------------------------------------------------------------
// Table entry n: (n, correctly rounded 1/n) for n = 1..8, (0, 0) otherwise.  The table is a STATIC
// shared array so its address is a compile-time constant: a lookup is shift + mask + one LDS.64.
__device__ __forceinline__ float2 div_entry(int code)
{
    const float c = (code >= 1 && code <= 8) ? (float)code : 0.0f;
    return make_float2(c, (c > 0.0f) ? __frcp_rn(c) : 0.0f);
}

__device__ __forceinline__ float2 div_pair(unsigned nib, const float2 *s_div)
{
    return s_div[nib];
}

// Fast quotient (exact for s == 0 and for every |s| >= 2^-100) plus the guard word of s:
// u = 2*bits(s) - 1 drops the sign and wraps +-0 to the top, so "0 < |s| < 2^-100" <=> u < GUARD_T.
constexpr unsigned GUARD_T = (27u << 24) - 1u;

// EXACT = true: the IEEE quotient itself (cr.x == 0 stands for an infinite divisor) -- a fused pass re-runs a
// chunk this way when one of its sums was a non-zero value below 2^-100, so its results equal IEEE division always.
// GUARD_Q (divisors 1..8 only, i.e. the binary kind): the guard word is taken from q0 = RN(s / c) instead of s.
// |q0| >= 2^-100 implies |s| >= 2^-100 (c >= 1), and q0 == 0 means s == 0, an infinite divisor (table entry (0, 0):
// the quotient is exactly 0 whatever s is) or a quotient that rounds to 0 either way -- so nothing is lost, and the
// non-conductive neighbours of a decaying isolated voxel (tiny sum, infinite divisor, exact result) raise no alarm.
template <bool EXACT = false, bool GUARD_Q = false>
__device__ __forceinline__ float div_fast(float s, float2 cr, unsigned &umin)
{
    if (EXACT) return __fdiv_rn(s, cr.x != 0.0f ? cr.x : __int_as_float(0x7f800000));
    const float q0 = __fmul_rn(s, cr.y);
    const float rem = __fmaf_rn(-q0, cr.x, s);
    umin = min(umin, __float_as_uint(GUARD_Q ? q0 : s) * 2u - 1u);
    return __fmaf_rn(rem, cr.y, q0);
}

// Neighbour sum in the reference's order (taufactor.py:97-102).
__device__ __forceinline__ float nbr_sum(float xp, float xm, float yp, float ym, float zp, float zm)
{
    float s = __fadd_rn(xp, xm);
    s = __fadd_rn(s, yp);
    s = __fadd_rn(s, ym);
    s = __fadd_rn(s, zp);
    return __fadd_rn(s, zm);
}

// f += omega * (q - f), taufactor.py:177-181, no FMA contraction.
__device__ __forceinline__ float relax(float c, float q, float omega)
{
    return __fadd_rn(c, __fmul_rn(__fsub_rn(q, c), omega));
}

// One binary-solver voxel update on the fast path; the caller checks umin once per group of updates.
template <bool EXACT = false>
__device__ __forceinline__ float sor_fast(float c, float xp, float xm, float yp, float ym, float zp, float zm,
                                          float2 cr, float omega, unsigned &umin)
{
    return relax(c, div_fast<EXACT, true>(nbr_sum(xp, xm, yp, ym, zp, zm), cr, umin), omega);
}

// The same update with a true IEEE division: taken only when some sum in the group is a non-zero
// value below 2^-100 (never seen in practice; keeps the result bit-identical for every finite input).
static __device__ __noinline__ float sor_exact(float c, float xp, float xm, float yp, float ym, float zp, float zm,
                                        float divisor, float omega)
{
    const float s = nbr_sum(xp, xm, yp, ym, zp, zm);
    const float q = (divisor > 0.0f) ? __fdiv_rn(s, divisor) : __fdiv_rn(s, __int_as_float(0x7f800000));
    return relax(c, q, omega);
}

// Colour updates of two voxels of a float4 group.  "xz" rows update components x and z (their z
// neighbours are y, w and the scalar zs = .w of the group on the left); "yw" rows update y and w
// (zs = .x of the group on the right).  xp/xm: x neighbours, up/dn: y+1 / y-1 neighbours.
// The *_fast forms return the new values in n0/n1 and fold the guard into umin; commit_* writes them
// after the (single, rare) exactness check of the caller.
__device__ __forceinline__ void xz_fast(const float4 &c, const float4 &xp, const float4 &xm, const float4 &up,
                                        const float4 &dn, float zs, unsigned code, const float2 *s_div, float omega,
                                        float &n0, float &n1, unsigned &umin)
{
    n0 = sor_fast(c.x, xp.x, xm.x, up.x, dn.x, c.y, zs, div_pair(code & 15u, s_div), omega, umin);
    n1 = sor_fast(c.z, xp.z, xm.z, up.z, dn.z, c.w, c.y, div_pair((code >> 8) & 15u, s_div), omega, umin);
}
__device__ __forceinline__ void yw_fast(const float4 &c, const float4 &xp, const float4 &xm, const float4 &up,
                                        const float4 &dn, float zs, unsigned code, const float2 *s_div, float omega,
                                        float &n0, float &n1, unsigned &umin)
{
    n0 = sor_fast(c.y, xp.y, xm.y, up.y, dn.y, c.z, c.x, div_pair((code >> 4) & 15u, s_div), omega, umin);
    n1 = sor_fast(c.w, xp.w, xm.w, up.w, dn.w, zs, c.z, div_pair((code >> 12) & 15u, s_div), omega, umin);
}
__device__ __forceinline__ void xz_exact(const float4 &c, const float4 &xp, const float4 &xm, const float4 &up,
                                         const float4 &dn, float zs, unsigned code, float omega, float &n0, float &n1)
{
    n0 = sor_exact(c.x, xp.x, xm.x, up.x, dn.x, c.y, zs, (float)(code & 15u), omega);
    n1 = sor_exact(c.z, xp.z, xm.z, up.z, dn.z, c.w, c.y, (float)((code >> 8) & 15u), omega);
}
__device__ __forceinline__ void yw_exact(const float4 &c, const float4 &xp, const float4 &xm, const float4 &up,
                                         const float4 &dn, float zs, unsigned code, float omega, float &n0, float &n1)
{
    n0 = sor_exact(c.y, xp.y, xm.y, up.y, dn.y, c.z, c.x, (float)((code >> 4) & 15u), omega);
    n1 = sor_exact(c.w, xp.w, xm.w, up.w, dn.w, zs, c.z, (float)((code >> 12) & 15u), omega);
}

// Single-row forms (generic kernel): fast path, one exactness check per row group.
__device__ __forceinline__ void update_xz(float4 &c, const float4 &xp, const float4 &xm, const float4 &up,
                                          const float4 &dn, float zs, unsigned code, const float2 *s_div, float omega)
{
    float n0, n1;
    unsigned umin = 0xffffffffu;
    xz_fast(c, xp, xm, up, dn, zs, code, s_div, omega, n0, n1, umin);
    if (umin < GUARD_T) xz_exact(c, xp, xm, up, dn, zs, code, omega, n0, n1);
    c.x = n0;
    c.z = n1;
}
__device__ __forceinline__ void update_yw(float4 &c, const float4 &xp, const float4 &xm, const float4 &up,
                                          const float4 &dn, float zs, unsigned code, const float2 *s_div, float omega)
{
    float n0, n1;
    unsigned umin = 0xffffffffu;
    yw_fast(c, xp, xm, up, dn, zs, code, s_div, omega, n0, n1, umin);
    if (umin < GUARD_T) yw_exact(c, xp, xm, up, dn, zs, code, omega, n0, n1);
    c.y = n0;
    c.w = n1;
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

// ------------------------------------------------------------------------------------------
// Multi-phase through stencil classes (TAUB_MULTIPHASE_CLASS).  Table row of class c (two float4):
//   {w_x+, w_x-, w_y+, w_y-}  {w_z+, w_z-, b, r}   with b = prefactor and r = RN(1/b) -- or b = r = 0
//   where the prefactor is inf (non-conductive voxel), which yields q = 0 without a special case.
// s as in taufactor.py:606-613 (each product rounded, summed left to right); q = s / b by
// q0 = RN(s*r), rem = s - q0*b (exact FMA), q = RN(q0 + rem*r): the fast path of IEEE division with the
// exactly rounded reciprocal (checked against s / b on 6e8 random pairs; same guard word as div_fast).
// The two half rows wa = {w_x+, w_x-, w_y+, w_y-}, wb = {w_z+, w_z-, b, r} are passed in: the fused kernel reads them
// from its shared-memory copy of the most frequent classes or, for the rare others, through the read-only path
// (two half-row arrays: a 32-byte sector holds the halves of two frequency-adjacent classes).
// ------------------------------------------------------------------------------------------
template <bool EXACT = false>
__device__ __forceinline__ float sor_class_rows(float c, float xp, float xm, float yp, float ym, float zp, float zm,
                                                const float4 &wa, const float4 &wb, float omega, unsigned &umin)
{
    float s = __fadd_rn(__fmul_rn(xp, wa.x), __fmul_rn(xm, wa.y));
    s = __fadd_rn(s, __fmul_rn(yp, wa.z));
    s = __fadd_rn(s, __fmul_rn(ym, wa.w));
    s = __fadd_rn(s, __fmul_rn(zp, wb.x));
    s = __fadd_rn(s, __fmul_rn(zm, wb.y));
    return relax(c, div_fast<EXACT>(s, make_float2(wb.z, wb.w), umin), omega);
}

// ------------------------------------------------------------------------------------------
// AnisotropicSolver (taufactor.py:422-478) through prefactor classes.  The prefactor, in the reference's
// fp32 accumulation order (:462-467), is ((((x- + x+) + Ky*y-) + Ky*y+) + Kz*z-) + Kz*z+ with the Dirichlet
// planes counting 2 -- it depends only on how many conductive neighbours a voxel has per axis, so at most
// 5 x 3 x 3 values occur.  codes = one uint16 class id per voxel (< TAUB_ANISO_CLASSES); lut = float2
// {b, RN(1/b)} per class (b = 1/b = 0: prefactor inf), then Ky, Kz.
// s = ((x+ + x-) + Ky*(y+ + y-)) + Kz*(z+ + z-)  (:475-477).
// ------------------------------------------------------------------------------------------
constexpr int ANISO_CLASSES = 64;

__device__ __forceinline__ float aniso_sum(float xp, float xm, float yp, float ym, float zp, float zm, float Ky, float Kz)
{
    float s = __fadd_rn(xp, xm);
    s = __fadd_rn(s, __fmul_rn(Ky, __fadd_rn(yp, ym)));
    return __fadd_rn(s, __fmul_rn(Kz, __fadd_rn(zp, zm)));
}

// generic kernel: IEEE division by the prefactor
__device__ __forceinline__ float sor_aniso(float c, float xp, float xm, float yp, float ym, float zp, float zm,
                                           float2 br, float Ky, float Kz, float omega)
{
    const float fac = br.x != 0.0f ? br.x : __int_as_float(0x7f800000);
    return relax(c, __fdiv_rn(aniso_sum(xp, xm, yp, ym, zp, zm, Ky, Kz), fac), omega);
}

// fused kernel: the exactly rounded reciprocal + one FMA correction (same fast path as div_fast)
template <bool EXACT = false>
__device__ __forceinline__ float sor_aniso_fast(float c, float xp, float xm, float yp, float ym, float zp, float zm,
                                                float2 br, float Ky, float Kz, float omega, unsigned &umin)
{
    return relax(c, div_fast<EXACT>(aniso_sum(xp, xm, yp, ym, zp, zm, Ky, Kz), br, umin), omega);
}

__device__ __forceinline__ int wrap(int a, int n)
{
    a %= n;
    return a < 0 ? a + n : a;
}

}  // namespace taub
