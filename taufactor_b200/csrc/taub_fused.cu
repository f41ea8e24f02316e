// taub_fused.cu -- temporally blocked two-colour sweep: TWO reference iterations per HBM pass.
//
// One pass reads the field once and writes it once while applying iteration t (colour A) and
// iteration t+1 (colour B) of taufactor.py:174-182, i.e. 4 B + 0.25 B of traffic per voxel per
// iteration instead of 8.5 (generic kernel) or 108 (reference eager path).
//
// Structure (per CTA): a (rows x z-groups) tile marched along x (the flux / slab axis).
//   * plane staging: every needed x-plane of the tile (output tile + 2 halo rows, + 1 halo float4
//     group each side) is brought into a 6-deep shared-memory ring by TMA: ONE
//     cp.async.bulk.tensor.3d box (rows x columns x 1 plane, out-of-bounds zero filled) per plane,
//     completion on an mbarrier, issued three planes ahead of use;
//     the 4-bit neighbour codes of the same box ride along in a second (uint16) TMA box;
//   * register rotation: each thread owns NI float4 columns and keeps a[p-2], a[p-1], raw[p],
//     raw[p+1] of its columns in registers, so x-neighbours never touch shared memory;
//   * wavefront: at step p colour A is applied to plane p (in place in shared memory -- legal
//     because a colour-A voxel only reads colour-B neighbours) and colour B to plane p-1, whose
//     result goes straight from registers to the destination buffer with 128-bit stores.
// y/z halos are recomputed by the neighbouring tile (overlapped tiling); the source buffer is
// read-only during the pass (ping-pong), so there is no inter-CTA hazard.  The arithmetic per
// voxel is the same correctly rounded sequence as the generic kernel: results are bit-identical.
#include <stdlib.h>

#include <cuda.h>   // CUtensorMap types only; the encoder is fetched through the runtime (no -lcuda)

#include "taub_common.cuh"

namespace taub {

constexpr int F_NT = 256;  // threads per CTA
constexpr int F_NB = 6;    // ring depth (planes in shared memory) of the binary kind

struct FusedParams {
    taub_geom g;
    const float *src;
    float *dst;
    const uint16_t *codes;   // binary: one uint16 (four 4-bit counts) per group; class kind: one uint16 per voxel
    const float *table;      // class kind: [2][n_classes][4] weight half-rows
    int n_classes;
    float omega;
    int colourA;
    int i_lo, i_hi;    // output planes (local)
    int a_lo, a_hi;    // planes that receive the colour-A update (output planes +-1, clipped to real planes)
    int LR, LG, LGc;   // loaded rows, box width in float4 groups (odd: see choose_tile), code-box width
    int LGt;           // groups per row that threads work on (= OG + 2 <= LG)
    int OR_, OG;       // output rows / groups per tile
    int tiles_k;
    int chunk_len;     // output planes per CTA
    const int *stop;   // optional device flag: non-zero -> the kernel returns at once
    float *peer_lo;    // optional: destination buffer of the rank below / above (peer memory); the first /
    float *peer_hi;    // last G output planes are also stored into its upper / lower ghost planes
    int slot_f4;       // float4 per field ring slot
    int cslot_h;       // uint16 per code ring slot
};

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, int count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity)
{
    uint32_t ok;
    do {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(ok)
            : "r"(bar), "r"(parity)
            : "memory");
    } while (!ok);
}
__device__ unsigned long long g_inexact_events = 0ULL;

// One TMA box: tensor coordinates (column, row, plane) in elements -> shared memory, signalling bar.
__device__ __forceinline__ void tma_load_3d(uint32_t dst, const CUtensorMap *map, int c0, int c1, int c2, uint32_t bar)
{
    asm volatile(
        "cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];" ::"r"(dst),
        "l"(reinterpret_cast<uint64_t>(map)), "r"(c0), "r"(c1), "r"(c2), "r"(bar)
        : "memory");
}

// 128-bit shared load that the compiler cannot split into scalar loads (a scalar load of one
// component of consecutive float4 groups is a 4-way bank conflict: same wavefronts, a quarter of the data).
__device__ __forceinline__ float4 lds128(const float4 *p)
{
    float4 v;
    asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(smem_u32(p)));
    return v;
}

// One row of a column: "xz" rows update components x and z, "yw" rows y and w.  zs is the one z
// neighbour that lives in the adjacent group (.w of the left group / .x of the right group).
__device__ __forceinline__ void row_update(const bool is_xz, float4 &c, const float4 &xp, const float4 &xm,
                                           const float4 &up, const float4 &dn, float zs, unsigned code,
                                           const float2 *s_div, float omega, unsigned &umin)
{
    float n0, n1;
    if (is_xz) {
        xz_fast(c, xp, xm, up, dn, zs, code, s_div, omega, n0, n1, umin);
        c.x = n0;
        c.z = n1;
    } else {
        yw_fast(c, xp, xm, up, dn, zs, code, s_div, omega, n0, n1, umin);
        c.y = n0;
        c.w = n1;
    }
}

// Class kind: cls2 = the row's four uint16 class ids (x | y << 16, z | w << 16).
__device__ __forceinline__ void row_update_class(const bool is_xz, float4 &c, const float4 &xp, const float4 &xm,
                                                 const float4 &up, const float4 &dn, float zs, uint2 cls2,
                                                 const float4 *tab, const float4 *tabB, float omega, unsigned &umin)
{
    if (is_xz) {
        const float n0 = sor_class(c.x, xp.x, xm.x, up.x, dn.x, c.y, zs, cls2.x & 0xffffu, tab, tabB, omega, umin);
        const float n1 = sor_class(c.z, xp.z, xm.z, up.z, dn.z, c.w, c.y, cls2.y & 0xffffu, tab, tabB, omega, umin);
        c.x = n0;
        c.z = n1;
    } else {
        const float n0 = sor_class(c.y, xp.y, xm.y, up.y, dn.y, c.z, c.x, cls2.x >> 16, tab, tabB, omega, umin);
        const float n1 = sor_class(c.w, xp.w, xm.w, up.w, dn.w, zs, c.z, cls2.y >> 16, tab, tabB, omega, umin);
        c.y = n0;
        c.w = n1;
    }
}

// Anisotropic kind: cls2 as above, the (b, 1/b) pair of a class from the static shared table.
__device__ __forceinline__ void row_update_aniso(const bool is_xz, float4 &c, const float4 &xp, const float4 &xm,
                                                 const float4 &up, const float4 &dn, float zs, uint2 cls2,
                                                 const float2 *s_div, float Ky, float Kz, float omega, unsigned &umin)
{
    if (is_xz) {
        const float n0 = sor_aniso_fast(c.x, xp.x, xm.x, up.x, dn.x, c.y, zs, s_div[cls2.x & 0xffffu], Ky, Kz, omega, umin);
        const float n1 = sor_aniso_fast(c.z, xp.z, xm.z, up.z, dn.z, c.w, c.y, s_div[cls2.y & 0xffffu], Ky, Kz, omega, umin);
        c.x = n0;
        c.z = n1;
    } else {
        const float n0 = sor_aniso_fast(c.y, xp.y, xm.y, up.y, dn.y, c.z, c.x, s_div[cls2.x >> 16], Ky, Kz, omega, umin);
        const float n1 = sor_aniso_fast(c.w, xp.w, xm.w, up.w, dn.w, zs, c.z, s_div[cls2.y >> 16], Ky, Kz, omega, umin);
        c.y = n0;
        c.w = n1;
    }
}

// The z neighbour from the adjacent lane's registers (one crossbar pass instead of a 4-way conflicted
// shared load); the two lanes at the warp ends read shared memory.  All 32 lanes must call this.
__device__ __forceinline__ float z_neighbour(const bool is_xz, const float4 &v, const float4 *buf, int i4, int lane,
                                             bool valid)
{
    float zs;
    if (is_xz) {
        zs = __shfl_up_sync(0xffffffffu, v.w, 1);
        if (lane == 0 && valid) zs = reinterpret_cast<const float *>(buf)[4 * i4 - 1];
    } else {
        zs = __shfl_down_sync(0xffffffffu, v.x, 1);
        if (lane == 31 && valid) zs = reinterpret_cast<const float *>(buf)[4 * i4 + 4];
    }
    return zs;
}

// Thread work item = a COLUMN of NRW vertically adjacent rows x one float4 group.  In every step the
// rows of a column alternate between "xz" and "yw" rows and swap roles each step; the column's internal
// y-neighbours stay in registers, only the rows above and below it come from shared memory.  PA0 = parity
// of the column's first row at step 0 (uniform over the whole grid, chosen by the host), so every step
// body is branch-free.
//
// OP ("odd periodic", experimental -- see taub_can_fuse): a periodic extent Ny or Nz is odd, so the wrap joins two
// voxels of the SAME colour and a ghost cell is the image of a voxel whose colour differs from the ghost's own
// index parity.  The reference reads ghost SNAPSHOTS taken before each iteration (taufactor.py:501-505); with
// the images loaded once per pass that is reproduced exactly by leaving the ghost ring of the odd axis out of
// the colour-A step: where the imaged voxel has colour B the snapshot before iteration t+1 equals the loaded
// value, and where it has colour A no colour-B voxel reads it.
template <int NRW, int PA0, int KIND, int NB, bool OP = false>   // KIND: TAUB_BINARY (4-bit codes) or TAUB_MULTIPHASE_CLASS (class ids)
__global__ void __launch_bounds__(F_NT, 2)      // NB: ring depth (planes of the tile resident in shared memory)
fused_sweep2_kernel(const FusedParams P, const __grid_constant__ CUtensorMap tmap,
                    const __grid_constant__ CUtensorMap cmap)
{
    // Programmatic dependent launch (opt-in, taub_iterate flags bit 1): let the next pass of the stream be
    // scheduled as soon as every CTA of this one has started, so that its launch latency and shared-memory
    // prologue overlap this pass's tail.  A no-op for an ordinary launch.
    pdl_trigger();
    extern __shared__ unsigned char smem_dyn[];
    // TMA destinations need 128-byte alignment: align the base, slots are multiples of 128 B
    unsigned char *smem_raw = smem_dyn + ((128u - (smem_u32(smem_dyn) & 127u)) & 127u);
    const taub_geom &g = P.g;
    constexpr bool ANI = (KIND == TAUB_ANISOTROPIC);
    constexpr bool CLS = (KIND == TAUB_MULTIPHASE_CLASS) || ANI;   // one uint16 id per voxel travels with the field
    constexpr int CPG = CLS ? 4 : 1;             // uint16 side-array elements per float4 group
    const int LR = P.LR, LG = P.LG, LGc = P.LGc;
    const float4 *tab = reinterpret_cast<const float4 *>(P.table);   // class kind: half rows A, then half rows B
    const float4 *tabB = tab + P.n_classes;
    const int plane_f4 = P.slot_f4;              // ring slot size in float4 (>= LR*LG, multiple of 8)
    const int cslot = P.cslot_h;
    float4 *planes = reinterpret_cast<float4 *>(smem_raw);
    uint16_t *cplanes = reinterpret_cast<uint16_t *>(smem_raw + (size_t)NB * plane_f4 * 16);
    uint64_t *mbar = reinterpret_cast<uint64_t *>(cplanes + (size_t)NB * cslot);
    __shared__ float2 s_div[ANISO_CLASSES];   // static: constant address, no address arithmetic per lookup

    const int tid = threadIdx.x, lane = tid & 31;
    const int tk = blockIdx.x % P.tiles_k, tj = blockIdx.x / P.tiles_k;
    const int b = blockIdx.z;
    const int c0 = P.i_lo + blockIdx.y * P.chunk_len;
    const int c1 = min(c0 + P.chunk_len, P.i_hi);
    const int R0 = tj * P.OR_, G0 = tk * P.OG;   // storage row / group of loaded (0, 0)
    const int PG = g.pitch >> 2;
    const int total_rel = c1 - c0 + 4;           // planes c0-2 .. c1+1
    const int64_t ps = g.plane_stride;

    if (!ANI && tid < ANISO_CLASSES) s_div[tid] = div_entry(tid);   // binary: (n, 1/n) of the neighbour count
    if (tid == 0) {
        for (int n = 0; n < NB; ++n) mbar_init(smem_u32(&mbar[n]), 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    // everything above touches only parameters and shared memory; from here on the kernel reads global memory
    // (stop flag, tables, source field): wait for the previous grid of the stream to complete and flush
    // (returns at once when this grid was not launched as a programmatic dependent)
    pdl_wait();
    if (P.stop && *P.stop) return;
    // anisotropic: (b, 1/b) of the prefactor classes; first used after the step loop's first __syncthreads
    if (ANI && tid < ANISO_CLASSES) s_div[tid] = reinterpret_cast<const float2 *>(P.table)[tid];
    const float Ky = ANI ? P.table[2 * ANISO_CLASSES] : 0.0f, Kz = ANI ? P.table[2 * ANISO_CLASSES + 1] : 0.0f;

    // one thread stages plane rel (local plane c0-2+rel) into ring slot rel % NB with one TMA box
    // (columns 4*G0.., rows R0.., one plane) plus the matching box of neighbour codes; the part of a
    // box outside the tensor reads as 0
    auto issue = [&](int rel) {
        const int slot = rel % NB;
        const uint32_t bar = smem_u32(&mbar[slot]);
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        mbar_expect_tx(bar, (uint32_t)(LR * (LG * 16 + LGc * CPG * 2)));
        const int pl = b * g.planes + (c0 - 2 + rel + G);
        tma_load_3d(smem_u32(planes + (size_t)slot * plane_f4), &tmap, 4 * G0, R0, pl, bar);
        tma_load_3d(smem_u32(cplanes + (size_t)slot * cslot), &cmap, G0 * CPG, R0, pl, bar);
    };
    if (tid == 0)
        for (int rel = 0; rel < min(NB - 1, total_rel); ++rel) issue(rel);

    // ---- this thread's column: rows lr0 .. lr0+NRW-1 of the tile, group gg
    const int NCT = (LR - 2) / NRW;          // columns stacked in the tile
    const int LGt = P.LGt;
    const int m = tid / LGt, gg = tid - m * LGt;
    const int lr0 = 1 + NRW * m;
    const int Ra = R0 + lr0, Gs = G0 + gg;
    const bool doit = (m < NCT) && (Gs < PG) && (Ra < g.rows);
    const bool colB = gg >= 1 && gg < LGt - 1 && Gs >= 1 && Gs < 1 + interior_groups(g.Nz);
    unsigned canB = 0;                       // bit r: row r of the column is an output row
#pragma unroll
    for (int r = 0; r < NRW; ++r)
        if (doit && colB && lr0 + r >= 2 && lr0 + r < LR - 2 && Ra + r >= G && Ra + r < G + g.Ny) canB |= 1u << r;
    const bool all_rows = (canB == (1u << NRW) - 1u);   // interior columns: every row is an output row
    // OP: ghost rows (odd Ny) keep their snapshot in the colour-A step, and so do the ghost columns k = -1
    // (.w of group 0) and k = Nz (component Nz % 4 of group (Nz + 4) / 4) for odd Nz
    unsigned keep_rows = 0;
    bool keep_lo_w = false, keep_hi_y = false, keep_hi_w = false;
    if (OP) {
        if (g.Ny & 1) {
#pragma unroll
            for (int r = 0; r < NRW; ++r)
                if (Ra + r < G || Ra + r >= G + g.Ny) keep_rows |= 1u << r;
        }
        if (g.Nz & 1) {
            keep_lo_w = (Gs == 0);
            keep_hi_y = (Gs == ((g.Nz + COL0) >> 2)) && ((g.Nz & 3) == 1);
            keep_hi_w = (Gs == ((g.Nz + COL0) >> 2)) && ((g.Nz & 3) == 3);
        }
    }
    const int i0 = lr0 * LG + gg;            // float4 index of row 0 inside a ring slot (row r: + r*LG)
    const int ic0 = (lr0 * LGc + gg) * CPG;  // uint16 index of row 0's code / class ids (row r: + r*LGc*CPG)
    // colour B first writes plane c0 (at step 2)
    float *dst0 = P.dst + (int64_t)b * g.image_stride + 4 * Gs + (int64_t)(c0 + G) * ps + (int64_t)Ra * g.pitch;

    // ---- register ring: rg[r][k] holds plane (c0-3+k+4j) of row r; at step s = 4j+ss:
    //      a[p-2] = rg[.][ss], a[p-1] = rg[.][ss+1], raw[p] -> a[p] = rg[.][ss+2], raw[p+1] = rg[.][ss+3]
    float4 rg[NRW][4];
    unsigned cr[NRW / 2][4];   // binary: neighbour codes, two rows per word, same ring positions
    mbar_wait(smem_u32(&mbar[0]), 0);
    mbar_wait(smem_u32(&mbar[1 % NB]), 0);
#pragma unroll
    for (int r = 0; r < NRW; ++r) {
#pragma unroll
        for (int k = 0; k < 4; ++k) rg[r][k] = make_float4(0.f, 0.f, 0.f, 0.f);
        if (doit) {
            rg[r][1] = planes[i0 + r * LG];
            rg[r][2] = planes[plane_f4 + i0 + r * LG];
        }
    }
#pragma unroll
    for (int q = 0; q < NRW / 2; ++q) {
#pragma unroll
        for (int k = 0; k < 4; ++k) cr[q][k] = 0;
        if (doit && !CLS)
            cr[q][2] = (unsigned)cplanes[cslot + ic0 + 2 * q * LGc] | ((unsigned)cplanes[cslot + ic0 + (2 * q + 1) * LGc] << 16);
    }

    unsigned umin = 0xffffffffu;   // guard word of every neighbour sum this thread divides
    const int n_steps = c1 - c0 + 2;
    for (int s4 = 0; s4 < n_steps; s4 += 4) {
#pragma unroll
        for (int ss = 0; ss < 4; ++ss) {
            const int s = s4 + ss;
            if (s >= n_steps) break;
            const int p = c0 - 1 + s;   // plane receiving colour A; colour B goes to plane p-1
            const int iM2 = ss & 3, iM1 = (ss + 1) & 3, iP = (ss + 2) & 3, iP1 = (ss + 3) & 3;
            __syncthreads();            // ring slot of plane p-2 is free; a[p-1] is visible in its slot
            if (tid == 0 && s - 1 + NB < total_rel) issue(s - 1 + NB);
            mbar_wait(smem_u32(&mbar[(s + 2) % NB]), (uint32_t)(((s + 2) / NB) & 1));
            const float4 *bufM1 = planes + (size_t)(s % NB) * plane_f4;
            float4 *bufP = planes + (size_t)((s + 1) % NB) * plane_f4;
            const float4 *bufP1 = planes + (size_t)((s + 2) % NB) * plane_f4;
            const uint16_t *codP1 = cplanes + (size_t)((s + 2) % NB) * cslot;
            const uint16_t *codP = cplanes + (size_t)((s + 1) % NB) * cslot;    // class kind reads ids in place
            const uint16_t *codM1 = cplanes + (size_t)(s % NB) * cslot;
            const bool doA = (p >= P.a_lo) && (p < P.a_hi);
            const bool keepA = (p >= c0) && (p < c1);   // a[p] is read by colour B of plane p next step
            const bool doB = (s >= 2);
            // one-sided halo exchange: output plane p-1 is one of the neighbour's ghost planes
            const bool send_lo = P.peer_lo != nullptr && (p - 1) < G;
            const bool send_hi = P.peer_hi != nullptr && (p - 1) >= g.Nx - G;
            if (doit) {
#pragma unroll
                for (int r = 0; r < NRW; ++r) rg[r][iP1] = bufP1[i0 + r * LG];
                if (!CLS) {
#pragma unroll
                    for (int q = 0; q < NRW / 2; ++q)
                        cr[q][iP1] = (unsigned)codP1[ic0 + 2 * q * LGc] | ((unsigned)codP1[ic0 + (2 * q + 1) * LGc] << 16);
                }
            }
            if (doA) {   // block-uniform
                float zs[NRW];
#pragma unroll
                for (int r = 0; r < NRW; ++r)
                    zs[r] = z_neighbour(((PA0 + ss + r) & 1) == 0, rg[r][iP], bufP, i0 + r * LG, lane, doit);
                if (doit) {
                    const float4 below = lds128(bufP + i0 - LG), above = lds128(bufP + i0 + NRW * LG);
#pragma unroll
                    for (int r = 0; r < NRW; ++r) {
                        // neighbours inside the column are registers; each row only reads the components
                        // its neighbours leave unchanged in this step
                        const float4 &dn = (r == 0) ? below : rg[r - 1][iP];
                        const float4 &up = (r == NRW - 1) ? above : rg[r + 1][iP];
                        const float4 snap = rg[r][iP];
                        if (ANI)
                            row_update_aniso(((PA0 + ss + r) & 1) == 0, rg[r][iP], rg[r][iP1], rg[r][iM1], up, dn, zs[r],
                                             *reinterpret_cast<const uint2 *>(codP + ic0 + r * LGc * CPG), s_div, Ky, Kz, P.omega, umin);
                        else if (CLS)
                            row_update_class(((PA0 + ss + r) & 1) == 0, rg[r][iP], rg[r][iP1], rg[r][iM1], up, dn, zs[r],
                                             *reinterpret_cast<const uint2 *>(codP + ic0 + r * LGc * CPG), tab, tabB, P.omega, umin);
                        else
                            row_update(((PA0 + ss + r) & 1) == 0, rg[r][iP], rg[r][iP1], rg[r][iM1], up, dn, zs[r],
                                       cr[r >> 1][iP] >> (16 * (r & 1)), s_div, P.omega, umin);
                        if (OP) {   // ghost cells of an odd periodic axis keep their snapshot
                            if (keep_rows & (1u << r)) {
                                rg[r][iP] = snap;
                            } else {
                                if (keep_lo_w || keep_hi_w) rg[r][iP].w = snap.w;
                                if (keep_hi_y) rg[r][iP].y = snap.y;
                            }
                        }
                    }
                    if (keepA) {
                        // other threads read the column's first and last row (their above / below) and,
                        // at the two ends of a warp, the neighbour lane's group (z_neighbour fall-back)
                        const bool edge_lane = (lane == 0) || (lane == 31);
#pragma unroll
                        for (int r = 0; r < NRW; ++r)
                            if (r == 0 || r == NRW - 1 || edge_lane) bufP[i0 + r * LG] = rg[r][iP];
                    }
                }
            }
            if (doB) {   // block-uniform
                float zs[NRW];
#pragma unroll
                for (int r = 0; r < NRW; ++r)
                    zs[r] = z_neighbour(((PA0 + ss + r) & 1) == 0, rg[r][iM1], bufM1, i0 + r * LG, lane, doit);
                if (canB) {
                    const float4 below = lds128(bufM1 + i0 - LG), above = lds128(bufM1 + i0 + NRW * LG);
#pragma unroll
                    for (int r = 0; r < NRW; ++r) {
                        const float4 &dn = (r == 0) ? below : rg[r - 1][iM1];
                        const float4 &up = (r == NRW - 1) ? above : rg[r + 1][iM1];
                        float4 out = rg[r][iM1];
                        if (ANI)
                            row_update_aniso(((PA0 + ss + r) & 1) == 0, out, rg[r][iP], rg[r][iM2], up, dn, zs[r],
                                             *reinterpret_cast<const uint2 *>(codM1 + ic0 + r * LGc * CPG), s_div, Ky, Kz, P.omega, umin);
                        else if (CLS)
                            row_update_class(((PA0 + ss + r) & 1) == 0, out, rg[r][iP], rg[r][iM2], up, dn, zs[r],
                                             *reinterpret_cast<const uint2 *>(codM1 + ic0 + r * LGc * CPG), tab, tabB, P.omega, umin);
                        else
                            row_update(((PA0 + ss + r) & 1) == 0, out, rg[r][iP], rg[r][iM2], up, dn, zs[r],
                                       cr[r >> 1][iM1] >> (16 * (r & 1)), s_div, P.omega, umin);
                        if (all_rows || (canB & (1u << r))) {
                            float *d = dst0 + (int64_t)r * g.pitch;
                            *reinterpret_cast<float4 *>(d) = out;
                            if (send_lo) *reinterpret_cast<float4 *>(P.peer_lo + (d - P.dst) + (int64_t)g.Nx * ps) = out;
                            if (send_hi) *reinterpret_cast<float4 *>(P.peer_hi + (d - P.dst) - (int64_t)g.Nx * ps) = out;
                        }
                    }
                }
                dst0 += ps;
            }
        }
    }
    // a non-zero sum below 2^-100 went through the fast division: count it (results stay within one
    // subnormal ulp of the reference there; never observed -- taub_inexact_events() reports it)
    if (umin < GUARD_T) atomicAdd(&g_inexact_events, 1ULL);
}

static size_t fused_smem_bytes(int LR, int LG, int LGc, int cpg, int nb)
{
    const size_t slot = ((size_t)(LR * LG + 7) / 8) * 8 * 16;               // fp32 box, 128-byte multiple
    const size_t cslot = ((size_t)(LR * LGc * cpg * 2 + 127) / 128) * 128;  // uint16 box (codes / class ids)
    return nb * (slot + cslot) + nb * 8 + 128;                    // + mbarriers, alignment slack (the division
                                                                      // table is 128 bytes of static shared memory)
}

constexpr int F_NRW = 4;   // rows per thread column

struct TileChoice {
    int LR, LG, LGc, LGt, OR_, OG, tiles_j, tiles_k;
    double eff;
};

// Tile = NCT columns (of F_NRW rows) stacked in y x LG float4 groups (OG = LG - 2 of them are outputs);
// one thread per (column, group).  TMA wants every box to start on a 16-byte boundary and to be a
// multiple of 16 bytes wide; for the uint16 code box that means OG (the tile step) is a multiple of 8
// groups and the code box is LG rounded up to 8.  A box is at most 256 elements wide (LG <= 64).  Pick
// the shape that wastes the fewest threads while two CTAs still fit in one SM's shared memory.
// Ring depth per kind.  The class kind carries 8 more bytes per float4 group in every slot and is bound by
// the L1 gather of its weight rows, not by HBM latency: a shallower ring (one plane of prefetch) buys
// ~1.6x larger tiles -> fewer halo re-loads and idle threads (measured: 6 -> 4 slots = +10 % at 384^3 / 512^3).
constexpr int F_NB_CLS = 4;
static int ring_depth(int cpg) { return cpg == 1 ? F_NB : F_NB_CLS; }

static TileChoice choose_tile(const taub_geom &g, int cpg)
{
    const int nb = ring_depth(cpg);
    // shared-memory budget per CTA (two CTAs per SM).  The class kind reads its weight rows through L1,
    // which shares the 256 KB with shared memory: leave it a little more.
    const size_t budget = (cpg == 1) ? 115000 : 106000;
    const int ng = interior_groups(g.Nz);
    TileChoice best{};
    best.eff = -1.0;
    for (int OG = 8; OG <= 56; OG += 8) {
        // threads work on LGt = OG + 2 groups per row.  Columns are F_NRW rows apart and 4*LG float4 is
        // a multiple of 32 banks when LG is even, so warps that straddle two columns bank-conflict; a
        // box one group wider (odd pitch) avoids that -- taken when it does not cost a column.
        const int LGt = OG + 2;
        int NCT = F_NT / LGt;   // columns the CTA's threads can cover
        while (NCT >= 1 && fused_smem_bytes(F_NRW * NCT + 2, LGt, ((LGt + 7) / 8) * 8, cpg, nb) > budget) --NCT;
        if (NCT < 1) continue;
        int LG = LGt;
        if (fused_smem_bytes(F_NRW * NCT + 2, LGt + 1, ((LGt + 8) / 8) * 8, cpg, nb) <= budget) LG = LGt + 1;
        const int LGc = ((LG + 7) / 8) * 8;
        const int NR = F_NRW * NCT, OR_ = NR - 2, LR = NR + 2;
        const int tj = ceil_div(g.Ny, OR_), tk = ceil_div(ng, OG);
        const double eff = ((double)g.Ny * ng) / ((double)tj * tk * F_NT * F_NRW);
        if (eff > best.eff + 1e-12) best = TileChoice{LR, LG, LGc, LGt, OR_, OG, tj, tk, eff};
        if (OG >= ng) break;          // one tile already spans the row
    }
    return best;
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *,
                                  const cuuint64_t *, const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn encode_tiled_fn()
{
    static EncodeTiledFn fn = nullptr;
    if (!fn) {
        void *ptr = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &q) == cudaSuccess &&
            q == cudaDriverEntryPointSuccess)
            fn = (EncodeTiledFn)ptr;
    }
    return fn;
}

// Encoding a tensor map costs a driver call; a solve alternates between the same few (buffer, box)
// combinations thousands of times, so keep the last few maps (per host thread).
struct MapKey {
    const void *base;
    int pitch, rows, planes_total, box_w, box_h, elem;
    bool operator==(const MapKey &o) const
    {
        return base == o.base && pitch == o.pitch && rows == o.rows && planes_total == o.planes_total &&
               box_w == o.box_w && box_h == o.box_h && elem == o.elem;
    }
};
struct MapCache {
    static constexpr int N = 8;
    MapKey key[N];
    CUtensorMap map[N];
    int used = 0, next = 0;
    const CUtensorMap *find(const MapKey &k) const
    {
        for (int i = 0; i < used; ++i)
            if (key[i] == k) return &map[i];
        return nullptr;
    }
    void put(const MapKey &k, const CUtensorMap &m)
    {
        key[next] = k;
        map[next] = m;
        next = (next + 1) % N;
        if (used < N) ++used;
    }
};
static thread_local MapCache g_maps;

// taub_iterate flags bit 1 (per host thread): launch the fused passes as programmatic dependents
thread_local bool g_fused_pdl = false;

// 3-D view of one ping-pong buffer: (columns = pitch, rows, bs * planes), fp32, box = LG*4 x LR x 1.
static int make_field_map(CUtensorMap *map, const taub_geom &g, const float *base, int LR, int LG)
{
    const MapKey key{base, g.pitch, g.rows, g.bs * g.planes, LG * 4, LR, 4};
    if (const CUtensorMap *hit = g_maps.find(key)) {
        *map = *hit;
        return TAUB_OK;
    }
    EncodeTiledFn enc = encode_tiled_fn();
    TAUB_REQUIRE(enc != nullptr, "cuTensorMapEncodeTiled is not available from this driver");
    const cuuint64_t dims[3] = {(cuuint64_t)g.pitch, (cuuint64_t)g.rows, (cuuint64_t)g.bs * g.planes};
    const cuuint64_t strides[2] = {(cuuint64_t)g.pitch * 4, (cuuint64_t)g.plane_stride * 4};
    const cuuint32_t box[3] = {(cuuint32_t)(LG * 4), (cuuint32_t)LR, 1};
    const cuuint32_t estr[3] = {1, 1, 1};
    const CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, const_cast<float *>(base), dims, strides, box, estr,
                           CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                           CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    TAUB_REQUIRE(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled failed with CUresult %d (pitch %d rows %d planes %lld box %dx%d)",
                 (int)r, g.pitch, g.rows, (long long)g.bs * g.planes, LG * 4, LR);
    g_maps.put(key, *map);
    return TAUB_OK;
}

// Same view of the uint16 side array (cpg elements per float4 group: 1 = neighbour codes, 4 = class ids):
// (pitch/4 * cpg, rows, bs * planes), box = LGc*cpg x LR x 1.
static int make_code_map(CUtensorMap *map, const taub_geom &g, const uint16_t *base, int LR, int LGc, int cpg)
{
    const MapKey key{base, g.pitch * cpg / 4, g.rows, g.bs * g.planes, LGc * cpg, LR, 2};
    if (const CUtensorMap *hit = g_maps.find(key)) {
        *map = *hit;
        return TAUB_OK;
    }
    EncodeTiledFn enc = encode_tiled_fn();
    TAUB_REQUIRE(enc != nullptr, "cuTensorMapEncodeTiled is not available from this driver");
    const cuuint64_t PG = (cuuint64_t)(g.pitch >> 2) * cpg;
    const cuuint64_t dims[3] = {PG, (cuuint64_t)g.rows, (cuuint64_t)g.bs * g.planes};
    const cuuint64_t strides[2] = {PG * 2, (cuuint64_t)g.rows * PG * 2};
    const cuuint32_t box[3] = {(cuuint32_t)(LGc * cpg), (cuuint32_t)LR, 1};
    const cuuint32_t estr[3] = {1, 1, 1};
    const CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_UINT16, 3, const_cast<uint16_t *>(base), dims, strides, box, estr,
                           CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                           CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    TAUB_REQUIRE(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled (codes) failed with CUresult %d", (int)r);
    g_maps.put(key, *map);
    return TAUB_OK;
}

// Plane chunks per tile column: fewest "waves x (planes + prologue)" on the resident-CTA capacity.
static void choose_chunks(int n_planes, int64_t tiles, int capacity, int *chunk_len, int *chunks)
{
    double best = 1e30;
    *chunk_len = n_planes + (n_planes & 1);
    *chunks = 1;
    for (int c = 1; c <= 512 && c <= (n_planes + 1) / 2; ++c) {
        int cl = ceil_div(n_planes, c);
        cl += cl & 1;   // even: every CTA starts with the same row parity
        const int ce = ceil_div(n_planes, cl);
        const int64_t waves = ceil_div64(tiles * ce, capacity);
        const double cost = (double)waves * (cl + 5);
        if (cost < best - 1e-9) {
            best = cost;
            *chunk_len = cl;
            *chunks = ce;
        }
    }
}

}  // namespace taub

using namespace taub;

static bool odd_periodic(const taub_geom &g) { return g.periodic && ((g.Ny & 1) || (g.Nz & 1)); }

static bool fuse_odd_periodic_enabled()
{
    static int on = -1;
    if (on < 0) {
        const char *e = getenv("TAUB_FUSE_ODD_PERIODIC");
        on = (e && e[0] == '1') ? 1 : 0;
    }
    return on == 1;
}

extern "C" {

unsigned long long taub_inexact_events(void)
{
    unsigned long long v = 0;
    if (cudaMemcpyFromSymbol(&v, g_inexact_events, sizeof(v)) != cudaSuccess) return ~0ULL;
    return v;
}

int taub_can_fuse(const taub_problem *p)
{
    if (!p || (p->kind != TAUB_BINARY && p->kind != TAUB_MULTIPHASE_CLASS && p->kind != TAUB_ANISOTROPIC) || !p->codes ||
        !p->field[0] || !p->field[1])
        return 0;
    if (p->kind != TAUB_BINARY && !p->lut) return 0;
    const taub_geom &g = p->g;
    // periodic wrap with odd Ny/Nz couples two voxels of the SAME colour (reference reads a ghost
    // snapshot): generic path, unless the experimental OP variant of the fused kernel (ghost ring of the odd
    // axis left out of the colour-A step) is switched on with TAUB_FUSE_ODD_PERIODIC=1.
    if (odd_periodic(g) && !fuse_odd_periodic_enabled()) return 0;
    if (g.bs > 65535) return 0;
    return choose_tile(g, (p->kind == TAUB_MULTIPHASE_CLASS || p->kind == TAUB_ANISOTROPIC) ? 4 : 1).eff > 0.0 ? 1 : 0;
}

int taub_fused_sweep2(const taub_problem *p, int64_t iter, int i_lo, int i_hi, void *stream)
{
    if (taub_can_fuse(p) != 1) {
        set_error("taub_fused_sweep2: problem does not qualify for the fused path");
        return TAUB_ERR_UNSUPPORTED;
    }
    const taub_geom &g = p->g;
    TAUB_REQUIRE(i_lo >= 0 && i_hi <= g.Nx && i_lo < i_hi, "taub_fused_sweep2: planes [%d, %d) outside the slab", i_lo, i_hi);
    const int cpg = (p->kind == TAUB_MULTIPHASE_CLASS || p->kind == TAUB_ANISOTROPIC) ? 4 : 1;
    const TileChoice t = choose_tile(g, cpg);
    FusedParams P;
    P.g = g;
    P.src = p->field[p->cur];
    P.dst = p->field[p->cur ^ 1];
    P.codes = p->codes;
    P.table = p->lut;
    P.n_classes = p->L;
    P.omega = p->omega;
    P.stop = p->stop;
    P.peer_lo = p->peer_lo[p->cur ^ 1];
    P.peer_hi = p->peer_hi[p->cur ^ 1];
    P.colourA = (int)(iter & 1);
    P.i_lo = i_lo;
    P.i_hi = i_hi;
    P.a_lo = max(i_lo - 1, -g.i_offset);
    P.a_hi = min(i_hi + 1, g.Nx_global - g.i_offset);
    P.LR = t.LR; P.LG = t.LG; P.LGc = t.LGc; P.LGt = t.LGt; P.OR_ = t.OR_; P.OG = t.OG;   // OR_ even, OG % 8 == 0
    P.tiles_k = t.tiles_k;
    const int n_planes = i_hi - i_lo;
    const int64_t tiles = (int64_t)t.tiles_j * t.tiles_k * g.bs;
    int dev_ord = 0;
    TAUB_CUDA(cudaGetDevice(&dev_ord));
    static int sm_count = 0;
    if (!sm_count) TAUB_CUDA(cudaDeviceGetAttribute(&sm_count, cudaDevAttrMultiProcessorCount, dev_ord));
    int chunk_len, chunks;
    choose_chunks(n_planes, tiles, 2 * sm_count, &chunk_len, &chunks);
    P.chunk_len = chunk_len;
    TAUB_REQUIRE(chunks <= 65535, "taub_fused_sweep2: too many plane chunks");
    P.slot_f4 = ((t.LR * t.LG + 7) / 8) * 8;
    P.cslot_h = ((t.LR * t.LGc * cpg * 2 + 127) / 128) * 64;
    const int nb = ring_depth(cpg);
    const size_t smem = fused_smem_bytes(t.LR, t.LG, t.LGc, cpg, nb);
    dim3 grid(t.tiles_j * t.tiles_k, chunks, g.bs);
    cudaStream_t s = (cudaStream_t)stream;
    CUtensorMap tmap, cmap;
    if (int rc = make_field_map(&tmap, g, P.src, t.LR, t.LG)) return rc;
    if (int rc = make_code_map(&cmap, g, P.codes, t.LR, t.LGc, cpg)) return rc;
    // parity of loaded row 1 (row a of every pair) at step 0, i.e. at plane c0-1: 0 -> x,z active.
    // Tile row offsets (multiples of the even OR_) and chunk starts (multiples of the even
    // chunk_len) do not change it, so it is one number for the whole grid.
    const int pa0 = (1 - G + g.i_offset + P.colourA + (i_lo - 1)) & 1;   // row lr = 1 is row 0 of a column
#define TAUB_LAUNCH_FUSED(PA_, KIND_, NB_, OP_)                                                                   \
    do {                                                                                                          \
        static size_t smem_set[64] = {0};   /* per device: raise the opt-in limit only when it grows */         \
        if (smem > smem_set[dev_ord & 63]) {                                                                      \
            TAUB_CUDA(cudaFuncSetAttribute(fused_sweep2_kernel<F_NRW, PA_, KIND_, NB_, OP_>,                      \
                                           cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));              \
            smem_set[dev_ord & 63] = smem;                                                                        \
        }                                                                                                         \
        TAUB_CUDA(launch_maybe_pdl(fused_sweep2_kernel<F_NRW, PA_, KIND_, NB_, OP_>, grid, dim3(F_NT), smem, s, P, \
                                   tmap, cmap));                                                                  \
    } while (0)
#define TAUB_LAUNCH_FUSED_PA(KIND_, NB_, OP_)                                                                     \
    do {                                                                                                          \
        if (pa0 == 0) TAUB_LAUNCH_FUSED(0, KIND_, NB_, OP_); else TAUB_LAUNCH_FUSED(1, KIND_, NB_, OP_);          \
    } while (0)
    const bool op = odd_periodic(g);
    if (p->kind == TAUB_MULTIPHASE_CLASS) {
        if (op) TAUB_LAUNCH_FUSED_PA(TAUB_MULTIPHASE_CLASS, F_NB_CLS, true); else TAUB_LAUNCH_FUSED_PA(TAUB_MULTIPHASE_CLASS, F_NB_CLS, false);
    } else if (p->kind == TAUB_ANISOTROPIC) {
        TAUB_LAUNCH_FUSED_PA(TAUB_ANISOTROPIC, F_NB_CLS, false);      // no periodic variant of this solver
    } else {
        if (op) TAUB_LAUNCH_FUSED_PA(TAUB_BINARY, F_NB, true); else TAUB_LAUNCH_FUSED_PA(TAUB_BINARY, F_NB, false);
    }
#undef TAUB_LAUNCH_FUSED_PA
#undef TAUB_LAUNCH_FUSED
    TAUB_CUDA(cudaGetLastError());
    count_launch();
    return TAUB_OK;
}

}  // extern "C"
