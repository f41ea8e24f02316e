// taub_sweep.cu -- generic one-colour sweep (any shape, both solver kinds), periodic ghost
// refresh, and the iteration driver.
//
// Replaces the body of the reference's hot loop, taufactor.py:174-182: nine to nineteen eager
// elementwise launches and 108-180 B/voxel of HBM traffic become ONE launch that reads the field
// once (4 B), the compressed prefactor (0.5 B nibble codes / 1 B phase index) and writes the
// field once (4 B) into the other ping-pong buffer.  Ping-pong makes the sweep a pure function
// of the source buffer, which is exactly the reference's snapshot semantics for periodic ghosts
// (taufactor.py:501-505 run BEFORE the update), including odd Ny/Nz where the wrap joins two
// voxels of the same colour.
#include <stdlib.h>

#include "taub_common.cuh"

namespace taub {

// ------------------------------------------------------------------------------------------
// Periodic ghost frame: rows [0,G) u [G+Ny, rows) over columns [2, Nz+6), and columns
// {2,3,Nz+4,Nz+5} of the interior rows, := image of the wrapped interior voxel.
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
refresh_ghosts_kernel(taub_geom g, float *__restrict__ f, int p_lo, const int *__restrict__ stop, int early_trigger)
{
    // early_trigger = 0 (taub_iterate flags bit 2): the sweep behind this refresh is released when the refresh
    // has finished instead of being parked on the SMs while it runs
    if (early_trigger) pdl_trigger();
    pdl_wait();      // before the first global read and before any thread exits (see launch_maybe_pdl)
    if (stop && *stop) return;
    // items: the 2G ghost rows as float4 groups (pitch/4 each), then for every interior row the left
    // and the right ghost column pair (one float2 each; columns 2,3 and Nz+4,Nz+5 are 8-byte aligned
    // when Nz is even, otherwise the pair is moved as two scalars).
    const int PG = g.pitch >> 2;
    const int n_row_items = 2 * G * PG;
    const int total = n_row_items + 2 * g.Ny;
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= total) return;
    float *plane = f + (int64_t)blockIdx.z * g.image_stride + (int64_t)(p_lo + blockIdx.y) * g.plane_stride;
    if (t < n_row_items) {
        const int r = t / PG, grp = t - r * PG;
        const int jr = r < G ? r : g.Ny + r;          // r in [G, 2G) -> rows [G+Ny, 2G+Ny)
        const int js = G + wrap(jr - G, g.Ny);
        float v[4];
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            const int c = 4 * grp + q;
            const bool frame = (c >= COL0 - G) && (c < COL0 + g.Nz + G);
            v[q] = frame ? plane[(int64_t)js * g.pitch + COL0 + wrap(c - COL0, g.Nz)] : 0.0f;
        }
        *reinterpret_cast<float4 *>(plane + (int64_t)jr * g.pitch + 4 * grp) = make_float4(v[0], v[1], v[2], v[3]);
    } else {
        const int u = t - n_row_items;
        const int r = u >> 1, side = u & 1;
        float *row = plane + (int64_t)(G + r) * g.pitch;
        const int c = side ? COL0 + g.Nz : COL0 - G;   // first of the two ghost columns
        const float v0 = row[COL0 + wrap(c - COL0, g.Nz)], v1 = row[COL0 + wrap(c + 1 - COL0, g.Nz)];
        if ((c & 1) == 0)
            *reinterpret_cast<float2 *>(row + c) = make_float2(v0, v1);
        else {
            row[c] = v0;
            row[c + 1] = v1;
        }
    }
}

// ------------------------------------------------------------------------------------------
// Generic half-sweep: one thread per interior float4 group (4 consecutive z voxels), the two
// voxels of the active colour are updated (warp-uniform branch on the row parity).  blockDim = (32, 8):
// x -> groups along z (coalesced 512 B per warp), y -> rows; grid.z -> (image, plane).
// ------------------------------------------------------------------------------------------
template <int KIND>   // 0 binary, 1 multi-phase (labels), 2 anisotropic, 3 multi-phase (stencil classes)
__global__ void __launch_bounds__(256)
half_sweep_kernel(taub_geom g, const float *__restrict__ src, float *__restrict__ dst,
                  const uint16_t *__restrict__ codes, const uint8_t *__restrict__ labels,
                  const float *__restrict__ lut, int L, float omega, int colour, int i_lo, int n_planes,
                  const int *__restrict__ stop)
{
    pdl_trigger();
    pdl_wait();      // before the first global read and before any thread exits (see launch_maybe_pdl)
    if (stop && *stop) return;
    constexpr bool MULTI = (KIND == TAUB_MULTIPHASE);
    __shared__ float2 s_div[16];
    extern __shared__ float s_lut[];  // MULTI: (L+1)^2
    const int tid = threadIdx.y * blockDim.x + threadIdx.x;
    if (!MULTI) {
        if (tid < 16) s_div[tid] = div_entry(tid);
    } else {
        for (int t = tid; t < (L + 1) * (L + 1); t += blockDim.x * blockDim.y) s_lut[t] = lut[t];
    }
    __syncthreads();

    const int ng = interior_groups(g.Nz);
    const int grp = blockIdx.x * blockDim.x + threadIdx.x;  // 0-based interior group
    const int j = blockIdx.y * blockDim.y + threadIdx.y;
    if (grp >= ng || j >= g.Ny) return;
    const int64_t ps = g.plane_stride;
    const int pitch = g.pitch;

    for (int z = blockIdx.z; z < g.bs * n_planes; z += gridDim.z) {
        const int b = z / n_planes;
        const int il = i_lo + (z - b * n_planes);  // local plane
        const int ig = il + g.i_offset;            // global plane
        const int64_t o = (int64_t)b * g.image_stride + (int64_t)(il + G) * ps +
                          (int64_t)(j + G) * pitch + COL0 + 4 * grp;
        float4 c4 = *reinterpret_cast<const float4 *>(src + o);
        const float4 xp4 = *reinterpret_cast<const float4 *>(src + o + ps);
        const float4 xm4 = *reinterpret_cast<const float4 *>(src + o - ps);
        const float4 yp4 = *reinterpret_cast<const float4 *>(src + o + pitch);
        const float4 ym4 = *reinterpret_cast<const float4 *>(src + o - pitch);
        // voxel q is active when (i + j + k) % 2 == colour; k = 4*grp + q so only q matters.
        // par0 is uniform over a warp (one row per warp): 0 -> x,z active, 1 -> y,w active.
        const int par0 = (ig + j + colour) & 1;
        if (KIND == TAUB_BINARY) {
            const unsigned code = codes[o >> 2];
            if (par0 == 0)
                update_xz(c4, xp4, xm4, yp4, ym4, src[o - 1], code, s_div, omega);
            else
                update_yw(c4, xp4, xm4, yp4, ym4, src[o + 4], code, s_div, omega);
        } else if (KIND == TAUB_MULTIPHASE_CLASS) {
            // codes: one uint16 class id per voxel; lut = one 8-float row {w_x+, w_x-, w_y+, w_y-, w_z+, w_z-, b, 1/b} per class
            // (b = 0 stands for an infinite prefactor); true IEEE division here, the fused kernel uses 1/b
            const uint2 cw = *reinterpret_cast<const uint2 *>(codes + o);
            const float4 *tab = reinterpret_cast<const float4 *>(lut);
#define TAUB_CLASS_Q(CLS, CEN, XP, XM, YP, YM, ZP, ZM)                                                     \
    {                                                                                                      \
        const float4 wa = __ldg(tab + 2 * (CLS)), wb = __ldg(tab + 2 * (CLS) + 1);                                   \
        float s = __fadd_rn(__fmul_rn(XP, wa.x), __fmul_rn(XM, wa.y));                                       \
        s = __fadd_rn(s, __fmul_rn(YP, wa.z));                                                              \
        s = __fadd_rn(s, __fmul_rn(YM, wa.w));                                                              \
        s = __fadd_rn(s, __fmul_rn(ZP, wb.x));                                                              \
        s = __fadd_rn(s, __fmul_rn(ZM, wb.y));                                                              \
        CEN = relax(CEN, __fdiv_rn(s, wb.z != 0.0f ? wb.z : __int_as_float(0x7f800000)), omega);            \
    }
            if (par0 == 0) {
                const float zl = src[o - 1];
                const float cy = c4.y;
                TAUB_CLASS_Q(cw.x & 0xffffu, c4.x, xp4.x, xm4.x, yp4.x, ym4.x, cy, zl)
                TAUB_CLASS_Q(cw.y & 0xffffu, c4.z, xp4.z, xm4.z, yp4.z, ym4.z, c4.w, cy)
            } else {
                const float zr = src[o + 4];
                const float cz = c4.z;
                TAUB_CLASS_Q(cw.x >> 16, c4.y, xp4.y, xm4.y, yp4.y, ym4.y, cz, c4.x)
                TAUB_CLASS_Q(cw.y >> 16, c4.w, xp4.w, xm4.w, yp4.w, ym4.w, zr, cz)
            }
#undef TAUB_CLASS_Q
        } else if (KIND == TAUB_ANISOTROPIC) {
            // codes: one uint16 prefactor-class id per voxel; lut = {b, 1/b} per class, then Ky, Kz
            const float2 *tab = reinterpret_cast<const float2 *>(lut);
            const float Ky = lut[2 * ANISO_CLASSES], Kz = lut[2 * ANISO_CLASSES + 1];
            const uint2 cw = *reinterpret_cast<const uint2 *>(codes + o);
            if (par0 == 0) {
                const float cy = c4.y;
                c4.x = sor_aniso(c4.x, xp4.x, xm4.x, yp4.x, ym4.x, cy, src[o - 1], __ldg(tab + (cw.x & 0xffffu)), Ky, Kz, omega);
                c4.z = sor_aniso(c4.z, xp4.z, xm4.z, yp4.z, ym4.z, c4.w, cy, __ldg(tab + (cw.y & 0xffffu)), Ky, Kz, omega);
            } else {
                const float cz = c4.z;
                c4.y = sor_aniso(c4.y, xp4.y, xm4.y, yp4.y, ym4.y, cz, c4.x, __ldg(tab + (cw.x >> 16)), Ky, Kz, omega);
                c4.w = sor_aniso(c4.w, xp4.w, xm4.w, yp4.w, ym4.w, src[o + 4], cz, __ldg(tab + (cw.y >> 16)), Ky, Kz, omega);
            }
        } else {
            const uint32_t lc = *reinterpret_cast<const uint32_t *>(labels + o);
            const uint32_t lxp = *reinterpret_cast<const uint32_t *>(labels + o + ps);
            const uint32_t lxm = *reinterpret_cast<const uint32_t *>(labels + o - ps);
            const uint32_t lyp = *reinterpret_cast<const uint32_t *>(labels + o + pitch);
            const uint32_t lym = *reinterpret_cast<const uint32_t *>(labels + o - pitch);
            const bool first = (ig == 0), last = (ig == g.Nx_global - 1);
            const int L1 = L + 1;
#define TAUB_MULTI_Q(Q, CEN, XP, XM, YP, YM, ZP, ZM, LZP, LZM)                                              \
    {                                                                                                       \
        const float *row = s_lut + ((lc >> (8 * Q)) & 255u) * L1;                                            \
        CEN = sor_multi(CEN, XP, XM, YP, YM, ZP, ZM, row[(lxp >> (8 * Q)) & 255u],                            \
                        row[(lxm >> (8 * Q)) & 255u], row[(lyp >> (8 * Q)) & 255u],                           \
                        row[(lym >> (8 * Q)) & 255u], row[LZP], row[LZM], first, last, omega);               \
    }
            if (par0 == 0) {
                const float zl = src[o - 1];
                const uint32_t lzl = labels[o - 1];
                const float cy = c4.y;
                TAUB_MULTI_Q(0, c4.x, xp4.x, xm4.x, yp4.x, ym4.x, cy, zl, (lc >> 8) & 255u, lzl)
                TAUB_MULTI_Q(2, c4.z, xp4.z, xm4.z, yp4.z, ym4.z, c4.w, cy, (lc >> 24) & 255u, (lc >> 8) & 255u)
            } else {
                const float zr = src[o + 4];
                const uint32_t lzr = labels[o + 4];
                const float cz = c4.z;
                TAUB_MULTI_Q(1, c4.y, xp4.y, xm4.y, yp4.y, ym4.y, cz, c4.x, (lc >> 16) & 255u, lc & 255u)
                TAUB_MULTI_Q(3, c4.w, xp4.w, xm4.w, yp4.w, ym4.w, zr, cz, lzr, (lc >> 16) & 255u)
            }
#undef TAUB_MULTI_Q
        }
        *reinterpret_cast<float4 *>(dst + o) = c4;
    }
}

}  // namespace taub

using namespace taub;

static thread_local bool g_refresh_late_trigger = false;   // taub_iterate flags bit 2

extern "C" {

static int refresh_ghosts(const taub_geom *g, float *field, int p_lo, int p_hi, const int *stop, void *stream);

int taub_refresh_ghosts(const taub_geom *g, float *field, int p_lo, int p_hi, void *stream)
{
    return refresh_ghosts(g, field, p_lo, p_hi, nullptr, stream);
}

static int refresh_ghosts(const taub_geom *g, float *field, int p_lo, int p_hi, const int *stop, void *stream)
{
    TAUB_REQUIRE(g && field, "taub_refresh_ghosts: null pointer");
    TAUB_REQUIRE(p_lo >= 0 && p_hi <= g->planes && p_lo < p_hi, "taub_refresh_ghosts: planes [%d, %d) invalid", p_lo, p_hi);
    const int total = 2 * G * (g->pitch >> 2) + 2 * g->Ny;
    for (int b0 = 0; b0 < g->bs; b0 += 65535) {
        dim3 grid(ceil_div(total, 256), p_hi - p_lo, min(g->bs - b0, 65535));
        TAUB_CUDA(launch_maybe_pdl(refresh_ghosts_kernel, grid, dim3(256), 0, (cudaStream_t)stream, *g,
                                   field + (int64_t)b0 * g->image_stride, p_lo, stop, g_refresh_late_trigger ? 0 : 1));
    }
    TAUB_CUDA(cudaGetLastError());
    count_launch();
    return TAUB_OK;
}

int taub_half_sweep(const taub_problem *p, int64_t iter, int i_lo, int i_hi, void *stream)
{
    TAUB_REQUIRE(p && p->field[0] && p->field[1], "taub_half_sweep: unbound problem");
    const taub_geom &g = p->g;
    TAUB_REQUIRE(i_lo >= -(G - 1) && i_hi <= g.Nx + (G - 1) && i_lo < i_hi,
                 "taub_half_sweep: planes [%d, %d) outside the slab", i_lo, i_hi);
    TAUB_REQUIRE(i_lo + g.i_offset >= 0 && i_hi + g.i_offset <= g.Nx_global,
                 "taub_half_sweep: planes [%d, %d) touch a Dirichlet plane", i_lo, i_hi);
    const float *src = p->field[p->cur];
    float *dst = p->field[p->cur ^ 1];
    const int colour = (int)(iter & 1);
    const int n_planes = i_hi - i_lo;
    dim3 block(32, 8);
    dim3 grid(ceil_div(interior_groups(g.Nz), 32), ceil_div(g.Ny, 8),
              (unsigned)min((int64_t)g.bs * n_planes, (int64_t)65535));
    TAUB_REQUIRE(grid.y <= 65535, "taub_half_sweep: Ny too large");
    cudaStream_t s = (cudaStream_t)stream;
    if (p->kind == TAUB_BINARY) {
        TAUB_REQUIRE(p->codes, "taub_half_sweep: binary problem without codes");
        TAUB_CUDA(launch_maybe_pdl(half_sweep_kernel<TAUB_BINARY>, grid, block, 0, s, g, src, dst, p->codes, nullptr, nullptr,
                                   0, p->omega, colour, i_lo, n_planes, p->stop));
    } else if (p->kind == TAUB_MULTIPHASE_CLASS) {
        TAUB_REQUIRE(p->codes && p->lut && p->L >= 1 && p->L <= 65536, "taub_half_sweep: class problem without table");
        TAUB_CUDA(launch_maybe_pdl(half_sweep_kernel<TAUB_MULTIPHASE_CLASS>, grid, block, 0, s, g, src, dst, p->codes, nullptr,
                                   p->lut, p->L, p->omega, colour, i_lo, n_planes, p->stop));
    } else if (p->kind == TAUB_ANISOTROPIC) {
        TAUB_REQUIRE(p->codes && p->lut, "taub_half_sweep: anisotropic problem without class ids / table");
        TAUB_REQUIRE(!g.periodic, "taub_half_sweep: the anisotropic solver has no periodic variant");
        TAUB_CUDA(launch_maybe_pdl(half_sweep_kernel<TAUB_ANISOTROPIC>, grid, block, 0, s, g, src, dst, p->codes, nullptr,
                                   p->lut, 0, p->omega, colour, i_lo, n_planes, p->stop));
    } else {
        TAUB_REQUIRE(p->labels && p->lut && p->L >= 1 && p->L <= TAUB_MAX_LABELS,
                     "taub_half_sweep: multi-phase problem without labels / table");
        const size_t smem = sizeof(float) * (p->L + 1) * (p->L + 1);
        TAUB_CUDA(launch_maybe_pdl(half_sweep_kernel<TAUB_MULTIPHASE>, grid, block, smem, s, g, src, dst, nullptr, p->labels,
                                   p->lut, p->L, p->omega, colour, i_lo, n_planes, p->stop));
    }
    TAUB_CUDA(cudaGetLastError());
    count_launch();
    return TAUB_OK;
}

int taub_iterate(taub_problem *p, int64_t iter, int n, int flags, void *stream)
{
    TAUB_REQUIRE(p && n >= 0, "taub_iterate: bad arguments");
    const taub_geom &g = p->g;
    TAUB_REQUIRE(g.i_offset == 0 && g.Nx == g.Nx_global,
                 "taub_iterate drives a whole volume; slabs interleave halo exchange in the caller");
    const bool fuse_ok = !(flags & 1) && taub_can_fuse(p) == 1;
    struct PdlScope {   // bit 1: the fused passes of THIS call are launched as programmatic dependents
        bool saved;
        explicit PdlScope(bool on) : saved(g_fused_pdl) { g_fused_pdl = on; }
        ~PdlScope() { g_fused_pdl = saved; }
    } pdl_scope((flags & 2) != 0);
    struct LateScope {
        bool saved;
        explicit LateScope(bool on) : saved(g_refresh_late_trigger) { g_refresh_late_trigger = on; }
        ~LateScope() { g_refresh_late_trigger = saved; }
    } late_scope((flags & 4) != 0);
    const bool reside_ok = !(flags & (1 | 8)) && taub_can_reside(p) == 1;
    int done = 0;
    while (done < n) {
        if (reside_ok && n - done >= 2) {
            // small volume: every remaining pair in one cooperative launch, field resident in shared memory
            const int pairs = (n - done) / 2;
            if (int rc = taub_resident_pairs(p, iter + done, pairs, stream)) return rc;
            done += 2 * pairs;
            continue;
        }
        if (g.periodic) {
            if (int rc = refresh_ghosts(&g, p->field[p->cur], 0, g.planes, p->stop, stream)) return rc;
        }
        if (fuse_ok && n - done >= 2) {
            if (int rc = taub_fused_sweep2(p, iter + done, 0, g.Nx, stream)) return rc;
            done += 2;
        } else {
            if (int rc = taub_half_sweep(p, iter + done, 0, g.Nx, stream)) return rc;
            done += 1;
        }
        p->cur ^= 1;
    }
    return TAUB_OK;
}

}  // extern "C"
