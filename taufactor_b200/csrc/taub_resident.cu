// taub_resident.cu -- small volumes: the field stays RESIDENT in shared memory for a whole block of iterations.
//
// The marching kernel of taub_fused.cu pays a pipeline fill of four planes per CTA and one launch per pass; below
// ~150^3 voxels that overhead is most of the time (100^3: 3.7 us per iteration however it is tiled).  Here ONE
// cooperative launch runs any number of iteration pairs of taufactor.py:174-182:
//   * the volume is cut into (x, y) bricks, one CTA per brick (<= one per SM), whole z rows; a brick and a 2-wide
//     frame of its neighbours' voxels live in shared memory for the whole launch;
//   * a pair = iteration t (colour A) on the brick plus the first ring of the frame, then iteration t+1 (colour B) on
//     the brick -- the same overlapped scheme as the fused kernel, so ONE exchange feeds TWO iterations;
//   * exchange through the two ping-pong field buffers in global memory (L2): after a pair a CTA stores the voxels
//     within 2 of its brick faces into the buffer of that pair and publishes a counter (release); before the next pair
//     it waits for its <= 8 neighbours' counters (acquire) and re-reads its frame.  Point-to-point, no grid barrier;
//     alternating buffers make the scheme race free (a CTA can run at most one pair ahead of a neighbour);
//   * shared-memory rows are stored COLOUR-SPLIT: the even columns of a row, then the odd ones.  The voxels one
//     iteration updates in a row are then contiguous: a work item is one float4 of four ACTIVE voxels, its x/y
//     neighbours are aligned float4s of the same half row, its z neighbours one float4 + one scalar of the other half.
//     No lane computes an inactive voxel, and E/O rows run the same instruction stream (selects, no branch).
// The arithmetic per voxel is the correctly rounded sequence of taub_common.cuh (reference order, no FMA, exact
// division: the fast path plus a per-item IEEE re-computation when a sum is a non-zero value below 2^-100), so the
// field is bit-identical to the other kernels and to the reference at every iteration.
//
// Not handled here (the marching kernels take over): periodic volumes with an odd Ny or Nz (a wrap that joins two
// voxels of one colour needs the snapshot rule of the fused kernel's OP variant), Nz < 8, slabs of a partitioned
// volume, bricks that do not fit 227 KB of shared memory.
#include <stdlib.h>

#include "taub_common.cuh"

namespace taub {

// threads per CTA: a template parameter of the kernel (512 or 1024; the steps of a pair are latency bound, so the
// second variant trades registers -- 64 per thread, smaller load batches -- for twice the warps)
constexpr size_t R_SMEM_MAX = 232448 - 1024;   // opt-in dynamic shared memory per CTA (227 KB) less the static part
constexpr int R_FLAG_STRIDE = 32;    // ints between two bricks' counters: one 128-byte line each
constexpr int R_MAX_BRICKS = 256;

struct ResParams {
    taub_geom g;
    float *buf[2];           // [0]: the current field at launch, [1]: the other ping-pong buffer
    const uint16_t *codes;   // binary kind: four 4-bit neighbour counts per float4 group; class kind: one id per voxel
    const float *table;      // class kind: [n_classes][8] weight rows (the last one inert)
    int n_classes, tab_k;    // ... rows in all / staged in shared memory
    float omega;
    int colour0;             // colour (iter & 1) of the first iteration
    int n_pairs;
    int nbx, nby;            // bricks per image along x and y
    int BX, BY;              // largest brick extents (shared-memory pitches)
    int ZS;                  // floats per shared-memory row: round_up(Nz + 8, 8); first half even, second half odd columns
    int *flags;              // one counter per brick (device, persistent): pairs completed, on top of epoch0
    int epoch0;
    int prof;                // accumulate the phase profile (TAUB_RESIDENT_PROF=1)
    const int *stop;
};

__device__ unsigned long long g_resident_timeouts = 0ULL;

// Counter protocol: the writer's bar.sync orders the CTA's stores before thread 0's fence + relaxed store; a reader
// polls with relaxed loads (no fence per poll) and fences once after it has seen the value.
__device__ __forceinline__ int ld_relaxed(const int *p)
{
    int v;
    asm volatile("ld.relaxed.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_relaxed(int *p, int v)
{
    asm volatile("st.relaxed.gpu.global.s32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ void fence_gpu() { asm volatile("fence.acq_rel.gpu;" ::: "memory"); }

// Cycle counts of the phases of a pair, summed over the pairs of the middle brick's thread 0 (taub_resident_profile):
// [0] wait for the neighbours' counters, [1] frame reload, [2] colour A, [3] colour B, [4] publish stores,
// [5] fence + counter store, [6] pairs, [7] whole launch.
__device__ unsigned long long g_resident_prof[8] = {0, 0, 0, 0, 0, 0, 0, 0};

// Brick extents: balanced cuts, sizes differ by at most one.
__host__ __device__ __forceinline__ int cut(int n, int parts, int i) { return (int)(((int64_t)n * i) / parts); }

struct Brick {
    int b, bi, bj;          // image, brick coordinates
    int x0, x1, y0, y1;     // owned voxels [x0, x1) x [y0, y1)
    int bx, by;
};

// ------------------------------------------------------------------------------------------------------
// Global <-> shared rows.  Shared row (li, lj) holds global x = x0 - 2 + li, y = y0 - 2 + lj (y wrapped for the
// periodic solvers), storage columns [0, ZS): half row E = even columns, half row O = odd columns.
// ------------------------------------------------------------------------------------------------------
__device__ __forceinline__ const float *global_row(const ResParams &P, const Brick &K, const float *base, int li, int lj)
{
    const taub_geom &g = P.g;
    const int gx = K.x0 - 2 + li, gy = K.y0 - 2 + lj;
    const int sr = g.periodic ? G + wrap(gy, g.Ny) : G + gy;
    return base + (int64_t)K.b * g.image_stride + (int64_t)(gx + G) * g.plane_stride + (int64_t)sr * g.pitch;
}

// Row tables (shared memory, built once per launch): everything a pair needs to know about a row is one 8-byte
// entry, so the steps of a pair do no index arithmetic (a pair is latency bound: ~1000 instructions per thread).
//   frame / publish rows: .x = float offset of the row in a field buffer (global), .y = float offset of the shared row
//                         (-1: skip -- a frame row outside the volume holds constants; publish: see RowTab::bnd)
//   colour-step rows:     .x = float offset of the shared row, .y = uint16 offset of its code row | parity << 30
struct RowTab {
    int2 *frame, *pub, *a_in, *a_out, *b;      // colour A: rows that read brick rows only / rows that need the frame
    int n_frame, n_pub, n_a_in, n_a_out, n_b;
};

// Rows f = warp, warp + R_WARPS, ... of a table from global memory into shared memory.  A warp first issues the loads
// of up to R_BATCH rows, then stores them: the latency is paid once per batch, not once per row.
template <int R_WARPS, int R_BATCH>      // R_BATCH: global loads a warp keeps in flight
__device__ __forceinline__ void load_rows(const int2 *tab, int n, float *fld, const float *base, int ZS, int warp, int lane)
{
    const int ZH = ZS >> 1, CG = ZS >> 2;
    for (int f0 = warp; f0 < n; f0 += R_WARPS * R_BATCH) {
        for (int g0 = 0; g0 < CG; g0 += 32) {
            const int g4 = g0 + lane;
            float4 v[R_BATCH];
            int dst[R_BATCH];
#pragma unroll
            for (int u = 0; u < R_BATCH; ++u) {
                const int f = f0 + u * R_WARPS;
                dst[u] = -1;
                if (f < n && g4 < CG) {
                    const int2 d = tab[f];
                    dst[u] = d.y;
                    if (d.y >= 0) v[u] = __ldcg(reinterpret_cast<const float4 *>(base + d.x) + g4);
                }
            }
#pragma unroll
            for (int u = 0; u < R_BATCH; ++u) {
                if (dst[u] >= 0) {
                    float *srow = fld + dst[u];
                    *reinterpret_cast<float2 *>(srow + 2 * g4) = make_float2(v[u].x, v[u].z);
                    *reinterpret_cast<float2 *>(srow + ZH + 2 * g4) = make_float2(v[u].y, v[u].w);
                }
            }
        }
    }
}

__device__ __forceinline__ void store_row(float *grow, const float *srow, int ZS, int Nz, int lane)
{
    const int ZH = ZS >> 1;
    // interior columns [4, Nz + 4): groups 1 .. (Nz + 3) / 4; a partial last group is stored voxel by voxel
    const int g_end = (Nz + COL0 + 3) >> 2;
    for (int g4 = 1 + lane; g4 < g_end; g4 += 32) {
        const float2 e = *reinterpret_cast<const float2 *>(srow + 2 * g4);
        const float2 o = *reinterpret_cast<const float2 *>(srow + ZH + 2 * g4);
        if (4 * g4 + 3 < Nz + COL0) {
            __stcg(reinterpret_cast<float4 *>(grow) + g4, make_float4(e.x, o.x, e.y, o.y));
        } else {
            const float v[4] = {e.x, o.x, e.y, o.y};
            for (int q = 0; q < 4; ++q)
                if (4 * g4 + q < Nz + COL0) __stcg(grow + 4 * g4 + q, v[q]);
        }
    }
}

// One colour step on the rows of a table (in place: an active voxel only reads voxels of the other colour).
// Thread = (row slot, float4 group of the active half row).
// Class kinds: the weight row of a class, from the shared copy of the first tab_k rows or -- rare classes -- global memory.
struct ResClassTab {
    const float4 *s, *g;    // shared copy / global table, two float4 per class
    unsigned k;             // rows staged
};
__device__ __forceinline__ void res_class_row(unsigned cls, const ResClassTab &T, float4 &wa, float4 &wb)
{
    if (cls < T.k) {
        wa = T.s[2 * cls];
        wb = T.s[2 * cls + 1];
    } else {
        wa = __ldg(T.g + 2 * cls);
        wb = __ldg(T.g + 2 * cls + 1);
    }
}

template <int KIND>
__device__ __forceinline__ void colour_step(const int2 *tab, int nrows, float *fld, const uint16_t *cod, const float2 *s_div,
                                            const ResClassTab &ctab, int colour, int ZS, int row_stride, float omega,
                                            int my_r, int my_q, int rows_per_round)
{
    const int ZH = ZS >> 1;
    if (my_r >= rows_per_round) return;
    for (int r = my_r; r < nrows; r += rows_per_round) {
        const int2 d = tab[r];
        // voxel (i, j, k) is active when (i + j + k) % 2 == colour; k and the storage column have the same parity
        const int par = ((d.y >> 30) ^ colour) & 1;                 // 0: even columns (E) active, 1: odd columns (O)
        float *row = fld + d.x;
        float *act = row + (par ? ZH : 0) + 4 * my_q;
        const float *oth = row + (par ? 0 : ZH) + 4 * my_q;
        float4 c = *reinterpret_cast<const float4 *>(act);
        const float4 xp = *reinterpret_cast<const float4 *>(act + row_stride);
        const float4 xm = *reinterpret_cast<const float4 *>(act - row_stride);
        const float4 yp = *reinterpret_cast<const float4 *>(act + ZS);
        const float4 ym = *reinterpret_cast<const float4 *>(act - ZS);
        const float4 z4 = *reinterpret_cast<const float4 *>(oth);
        const float ze = par ? oth[4] : oth[-1];
        // E active: z+ = O[m], z- = O[m-1];  O active: z- = E[m], z+ = E[m+1]
        const float4 zp = par ? make_float4(z4.y, z4.z, z4.w, ze) : z4;
        const float4 zm = par ? z4 : make_float4(ze, z4.x, z4.y, z4.z);
        if (KIND == TAUB_MULTIPHASE_CLASS) {
            // class ids of the row are stored colour-split like the field: four consecutive uint16 of the active half
            const uint2 iw = *reinterpret_cast<const uint2 *>(cod + (d.y & 0x3fffffff) + (par ? ZH : 0) + 4 * my_q);
            float4 wa0, wb0, wa1, wb1, wa2, wb2, wa3, wb3;
            res_class_row(iw.x & 0xffffu, ctab, wa0, wb0);
            res_class_row(iw.x >> 16, ctab, wa1, wb1);
            res_class_row(iw.y & 0xffffu, ctab, wa2, wb2);
            res_class_row(iw.y >> 16, ctab, wa3, wb3);
            unsigned um = 0xffffffffu;
            float n0 = sor_class_rows(c.x, xp.x, xm.x, yp.x, ym.x, zp.x, zm.x, wa0, wb0, omega, um);
            float n1 = sor_class_rows(c.y, xp.y, xm.y, yp.y, ym.y, zp.y, zm.y, wa1, wb1, omega, um);
            float n2 = sor_class_rows(c.z, xp.z, xm.z, yp.z, ym.z, zp.z, zm.z, wa2, wb2, omega, um);
            float n3 = sor_class_rows(c.w, xp.w, xm.w, yp.w, ym.w, zp.w, zm.w, wa3, wb3, omega, um);
            if (um < GUARD_T) {   // a non-zero sum below 2^-100: IEEE division, never the fast path
                n0 = sor_class_rows<true>(c.x, xp.x, xm.x, yp.x, ym.x, zp.x, zm.x, wa0, wb0, omega, um);
                n1 = sor_class_rows<true>(c.y, xp.y, xm.y, yp.y, ym.y, zp.y, zm.y, wa1, wb1, omega, um);
                n2 = sor_class_rows<true>(c.z, xp.z, xm.z, yp.z, ym.z, zp.z, zm.z, wa2, wb2, omega, um);
                n3 = sor_class_rows<true>(c.w, xp.w, xm.w, yp.w, ym.w, zp.w, zm.w, wa3, wb3, omega, um);
            }
            *reinterpret_cast<float4 *>(act) = make_float4(n0, n1, n2, n3);
            continue;
        }
        // neighbour counts: code words of storage groups 2q, 2q+1; E voxels are nibbles 0 and 2, O voxels 1 and 3
        const unsigned cw = *reinterpret_cast<const unsigned *>(cod + (d.y & 0x3fffffff) + 2 * my_q) >> (par ? 4 : 0);
        const float2 d0 = s_div[cw & 15u], d1 = s_div[(cw >> 8) & 15u], d2 = s_div[(cw >> 16) & 15u],
                     d3 = s_div[(cw >> 24) & 15u];
        unsigned um = 0xffffffffu;
        float n0 = sor_fast(c.x, xp.x, xm.x, yp.x, ym.x, zp.x, zm.x, d0, omega, um);
        float n1 = sor_fast(c.y, xp.y, xm.y, yp.y, ym.y, zp.y, zm.y, d1, omega, um);
        float n2 = sor_fast(c.z, xp.z, xm.z, yp.z, ym.z, zp.z, zm.z, d2, omega, um);
        float n3 = sor_fast(c.w, xp.w, xm.w, yp.w, ym.w, zp.w, zm.w, d3, omega, um);
        if (um < GUARD_T) {   // a non-zero sum below 2^-100: IEEE division (taufactor.py:177), never the fast path
            n0 = sor_exact(c.x, xp.x, xm.x, yp.x, ym.x, zp.x, zm.x, d0.x, omega);
            n1 = sor_exact(c.y, xp.y, xm.y, yp.y, ym.y, zp.y, zm.y, d1.x, omega);
            n2 = sor_exact(c.z, xp.z, xm.z, yp.z, ym.z, zp.z, zm.z, d2.x, omega);
            n3 = sor_exact(c.w, xp.w, xm.w, yp.w, ym.w, zp.w, zm.w, d3.x, omega);
        }
        *reinterpret_cast<float4 *>(act) = make_float4(n0, n1, n2, n3);
    }
}

template <int KIND, int R_NT, int R_BATCH>
__global__ void __launch_bounds__(R_NT, 1)
resident_kernel(const ResParams P)
{
    constexpr int R_WARPS = R_NT / 32;
    constexpr bool CLS = (KIND == TAUB_MULTIPHASE_CLASS);
    extern __shared__ __align__(16) unsigned char r_smem[];
    __shared__ __align__(128) float2 s_div[16];
    const taub_geom &g = P.g;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (P.stop && *P.stop) return;      // uniform over the grid: set by a check queued before this launch
    if (tid < 16) s_div[tid] = div_entry(tid);

    Brick K;
    {
        int id = blockIdx.x;
        K.bj = id % P.nby;
        id /= P.nby;
        K.bi = id % P.nbx;
        K.b = id / P.nbx;
        K.x0 = cut(g.Nx, P.nbx, K.bi);
        K.x1 = cut(g.Nx, P.nbx, K.bi + 1);
        K.y0 = cut(g.Ny, P.nby, K.bj);
        K.y1 = cut(g.Ny, P.nby, K.bj + 1);
        K.bx = K.x1 - K.x0;
        K.by = K.y1 - K.y0;
    }
    const bool per = g.periodic != 0;
    const int ZS = P.ZS, ZH = ZS >> 1, RY = P.BY + 4, CY = P.BY + 2, CG = ZS >> 2;
    const int LX = K.bx + 4, LY = K.by + 4;     // rows held: the brick and its 2-wide frame
    float *fld = reinterpret_cast<float *>(r_smem);                                 // [BX+4][BY+4][ZS]
    // binary: neighbour codes [BX+2][BY+2][ZS/4]; class kinds: ids [BX+2][BY+2][ZS] (colour-split rows)
    uint16_t *cod = reinterpret_cast<uint16_t *>(fld + (size_t)(P.BX + 4) * RY * ZS);
    const int CROW = CLS ? ZS : CG;             // uint16 per code / id row
    float4 *s_tab = reinterpret_cast<float4 *>(cod + (((size_t)(P.BX + 2) * CY * CROW + 7) & ~(size_t)7));   // 16-byte aligned
    RowTab T;
    T.frame = reinterpret_cast<int2 *>(s_tab + 2 * (CLS ? P.tab_k : 0));
    T.pub = T.frame + 4 * (P.BY + 4) + 4 * P.BX;
    T.a_out = T.pub + P.BX * P.BY;
    T.b = T.a_out + (P.BX + 2) * (P.BY + 2);
    T.a_in = T.b + P.BX * P.BY;
    // the steps never touch a Dirichlet plane, nor (no-flux solvers) a row outside the volume
    const int a_li0 = (K.x0 == 0) ? 2 : 1, a_li1 = (K.x1 == g.Nx) ? K.bx + 2 : K.bx + 3;
    const int a_lj0 = (!per && K.y0 == 0) ? 2 : 1, a_lj1 = (!per && K.y1 == g.Ny) ? K.by + 2 : K.by + 3;
    T.n_frame = 4 * LY + 4 * K.bx;
    T.n_pub = K.bx * K.by;
    // colour A on a brick row at least one row inside the brick reads brick rows only: it does not wait for the frame
    const int in_x = max(K.bx - 2, 0), in_y = max(K.by - 2, 0);
    T.n_a_in = in_x * in_y;
    T.n_a_out = (a_li1 - a_li0) * (a_lj1 - a_lj0) - T.n_a_in;
    T.n_b = K.bx * K.by;
    // global float offset of shared row (li, lj): x = x0 - 2 + li, y = y0 - 2 + lj (wrapped for the periodic solvers)
    auto goff = [&](int li, int lj) {
        const int gx = K.x0 - 2 + li, gy = K.y0 - 2 + lj;
        const int sr = per ? G + wrap(gy, g.Ny) : G + gy;
        return (int)((int64_t)K.b * g.image_stride + (int64_t)(gx + G) * g.plane_stride + (int64_t)sr * g.pitch);
    };
    auto soff = [&](int li, int lj) { return (li * RY + lj) * ZS; };
    for (int f = tid; f < T.n_frame; f += R_NT) {
        // the frame: two bands of 2 x LY rows below / above the brick, then 4 rows beside each of its bx planes
        int li, lj;
        if (f < 4 * LY) {
            const int band = f / LY;                  // 0, 1: planes 0, 1; 2, 3: planes bx+2, bx+3
            li = band < 2 ? band : K.bx + band;
            lj = f - band * LY;
        } else {
            const int e = f - 4 * LY, c = e & 3;
            li = 2 + (e >> 2);
            lj = c < 2 ? c : K.by + c;
        }
        const int gx = K.x0 - 2 + li, gy = K.y0 - 2 + lj;
        const bool moving = gx >= 0 && gx < g.Nx && (per || (gy >= 0 && gy < g.Ny));   // outside: constants, loaded once
        T.frame[f] = make_int2(goff(li, lj), moving ? soff(li, lj) : -1);
    }
    for (int r = tid; r < T.n_pub; r += R_NT) {
        const int oi = r / K.by, oj = r - oi * K.by;
        const bool bnd = oi < 2 || oi >= K.bx - 2 || oj < 2 || oj >= K.by - 2;   // a neighbour's frame covers this row
        // .y: shared offset, bit 30 = interior row (published after the last pair only)
        T.pub[r] = make_int2(goff(oi + 2, oj + 2), soff(oi + 2, oj + 2) | (bnd ? 0 : 1 << 30));
        T.b[r] = make_int2(soff(oi + 2, oj + 2), (((oi + 1) * CY + oj + 1) * CROW) | (((K.x0 + oi + K.y0 + oj) & 1) << 30));
    }
    auto a_entry = [&](int li, int lj) {
        return make_int2(soff(li, lj), (((li - 1) * CY + lj - 1) * CROW) | (((K.x0 + li + K.y0 + lj) & 1) << 30));
    };
    for (int r = tid; r < T.n_a_in; r += R_NT) T.a_in[r] = a_entry(3 + r / in_y, 3 + r % in_y);
    if (tid == 0) {          // the rest of the colour-A region, in order (a few hundred rows, once per launch)
        int m = 0;
        for (int li = a_li0; li < a_li1; ++li)
            for (int lj = a_lj0; lj < a_lj1; ++lj)
                if (!(li >= 3 && li < K.bx + 1 && lj >= 3 && lj < K.by + 1)) T.a_out[m++] = a_entry(li, lj);
    }

    // ---- start: the brick and its frame from the current field, neighbour codes of the brick and its first ring
    for (int r = warp; r < LX * LY; r += R_WARPS) {
        const int li = r / LY, lj = r - li * LY;
        const float *grow = P.buf[0] + goff(li, lj);
        float *srow = fld + soff(li, lj);
        for (int g4 = lane; g4 < CG; g4 += 32) {
            const float4 v = __ldcg(reinterpret_cast<const float4 *>(grow) + g4);
            *reinterpret_cast<float2 *>(srow + 2 * g4) = make_float2(v.x, v.z);
            *reinterpret_cast<float2 *>(srow + ZH + 2 * g4) = make_float2(v.y, v.w);
        }
    }
    for (int r = warp; r < (LX - 2) * (LY - 2); r += R_WARPS) {
        const int ci = r / (LY - 2), cj = r - ci * (LY - 2);
        const int gx = K.x0 - 1 + ci, gy = K.y0 - 1 + cj;
        const bool inside = gx >= 0 && gx < g.Nx && (per || (gy >= 0 && gy < g.Ny));
        const int sr = per ? G + wrap(gy, g.Ny) : G + gy;
        if (CLS) {
            // one uint16 class id per storage voxel; rows outside the volume: the inert class (the table's last row)
            const uint16_t *irow = P.codes + (((int64_t)K.b * g.planes + (gx + G)) * g.rows + sr) * g.pitch;
            uint16_t *srow = cod + (size_t)(ci * CY + cj) * ZS;
            const unsigned inert = (unsigned)(P.n_classes - 1);
            for (int g4 = lane; g4 < CG; g4 += 32) {
                uint2 w = make_uint2(inert | (inert << 16), inert | (inert << 16));
                if (inside) w = __ldg(reinterpret_cast<const uint2 *>(irow) + g4);      // ids of columns 4g4 .. 4g4+3
                *reinterpret_cast<unsigned *>(srow + 2 * g4) = (w.x & 0xffffu) | (w.y << 16);            // even columns
                *reinterpret_cast<unsigned *>(srow + ZH + 2 * g4) = (w.x >> 16) | (w.y & 0xffff0000u);   // odd columns
            }
            continue;
        }
        const uint16_t *grow = P.codes + ((int64_t)K.b * g.planes + (gx + G)) * g.rows * (g.pitch >> 2) + (int64_t)sr * (g.pitch >> 2);
        for (int g4 = lane; g4 < CG; g4 += 32) {
            unsigned w = inside ? (unsigned)__ldg(grow + g4) : 0u;
            // only interior columns [4, Nz + 4) are ever updated: ghost / padding columns get count 0
            unsigned keep = 0;
            for (int q = 0; q < 4; ++q)
                if (4 * g4 + q >= COL0 && 4 * g4 + q < g.Nz + COL0) keep |= 0xfu << (4 * q);
            cod[(size_t)(ci * CY + cj) * CG + g4] = (uint16_t)(w & keep);
        }
    }

    if (CLS)
        for (int t = tid; t < 2 * P.tab_k; t += R_NT) s_tab[t] = __ldg(reinterpret_cast<const float4 *>(P.table) + t);
    const ResClassTab ctab{s_tab, reinterpret_cast<const float4 *>(P.table), (unsigned)(CLS ? P.tab_k : 0)};
    const int QN = ZS >> 3;                         // float4 groups per half row
    const int rows_per_round = R_NT / QN;
    const int my_r = tid / QN, my_q = tid - my_r * QN;
    const int row_stride = RY * ZS;

    // neighbour bricks whose counters gate this brick's frame (threads 0..8, centre excluded); every counter has a
    // 128-byte line of its own: hundreds of pollers on one line would delay the very store they are waiting for
    int nb_flag = -1;
    if (tid < 9 && tid != 4) {
        const int nbi = K.bi + tid / 3 - 1;
        int nbj = K.bj + tid % 3 - 1;
        bool ok = nbi >= 0 && nbi < P.nbx;
        if (per)
            nbj = wrap(nbj, P.nby);
        else
            ok = ok && nbj >= 0 && nbj < P.nby;
        if (ok && !(nbi == K.bi && nbj == K.bj)) nb_flag = ((K.b * P.nbx + nbi) * P.nby + nbj) * R_FLAG_STRIDE;
    }
    // periodic z: ghost column 3 := column Nz + 3 (odd), ghost column Nz + 4 := column 4 (even); Nz is even here
    auto z_ghosts = [&](const int2 *tab, int n) {
        for (int r = tid; r < n; r += R_NT) {
            float *row = fld + tab[r].x;
            row[ZH + 1] = row[ZH + ((g.Nz + 3) >> 1)];
            row[(g.Nz + COL0) >> 1] = row[COL0 >> 1];
        }
    };
    // colour A on the inner rows runs on warps 1.. while warp 0 polls the neighbours' counters
    const int rows_per_round_w = (R_NT - 32) / QN;
    const int my_r_w = tid >= 32 ? (tid - 32) / QN : rows_per_round_w, my_q_w = tid >= 32 ? (tid - 32) % QN : 0;
    if (per) z_ghosts(T.b, T.n_b);      // (the brick's own rows; visible after the first barrier of the pair loop)
    __syncthreads();

    const bool prof = (P.prof != 0 && tid == 0 && blockIdx.x == gridDim.x / 2);
    const long long t_launch = clock64();
    long long t_prev = t_launch;
#define PROF(slot)                                                                         \
    if (prof) {                                                                            \
        const long long t_now = clock64();                                                 \
        atomicAdd(&g_resident_prof[slot], (unsigned long long)(t_now - t_prev));           \
        t_prev = t_now;                                                                    \
    }
    for (int n = 0; n < P.n_pairs; ++n) {
        float *wbuf = P.buf[(n & 1) ^ 1];            // pair n publishes into buf[1], buf[0], buf[1], ...
        if (prof) t_prev = clock64();
        // ---- colour A on the rows that do not need the frame (warps 1..), under the wait for the neighbours
        colour_step<KIND>(T.a_in, T.n_a_in, fld, cod, s_div, ctab, P.colour0, ZS, row_stride, P.omega, my_r_w, my_q_w, rows_per_round_w);
        if (n > 0) {
            // ---- wait for the neighbours' pair n-1, then re-read the frame from the buffer they wrote
            if (nb_flag >= 0) {
                const int target = P.epoch0 + n;
                const long long t0 = clock64();
                while (ld_relaxed(P.flags + nb_flag) - target < 0) {
                    __nanosleep(20);
                    if (clock64() - t0 > 4000000000LL) {     // ~2 s: never in a correct run; do not hang the device
                        atomicAdd(&g_resident_timeouts, 1ULL);
                        break;
                    }
                }
                fence_gpu();
            }
            __syncthreads();
            PROF(0);
            load_rows<R_WARPS, R_BATCH>(T.frame, T.n_frame, fld, P.buf[((n - 1) & 1) ^ 1], ZS, warp, lane);
        }
        __syncthreads();
        PROF(1);
        if (per) {
            z_ghosts(T.a_out, T.n_a_out);      // ring rows just loaded (and, again, the brick's edge rows)
            __syncthreads();
        }
        colour_step<KIND>(T.a_out, T.n_a_out, fld, cod, s_div, ctab, P.colour0, ZS, row_stride, P.omega, my_r, my_q, rows_per_round);
        __syncthreads();
        PROF(2);
        if (per) {
            z_ghosts(T.b, T.n_b);
            __syncthreads();
        }
        colour_step<KIND>(T.b, T.n_b, fld, cod, s_div, ctab, P.colour0 ^ 1, ZS, row_stride, P.omega, my_r, my_q, rows_per_round);
        __syncthreads();
        if (per) z_ghosts(T.b, T.n_b);          // for the next pair's early colour-A rows (barriers below come first)
        PROF(3);
        // ---- publish: the voxels the neighbours' frames cover (everything after the last pair)
        const int skip_mask = (n == P.n_pairs - 1) ? 0 : 1 << 30;
        for (int r = warp; r < T.n_pub; r += R_WARPS) {
            const int2 d = T.pub[r];
            if (!(d.y & skip_mask)) store_row(wbuf + d.x, fld + (d.y & 0x3fffffff), ZS, g.Nz, lane);
        }
        __syncthreads();
        PROF(4);
        // bar.sync orders the CTA's stores before thread 0's fence, which is cumulative: a neighbour that reads the
        // counter and fences sees every row published above
        if (tid == 0) {
            fence_gpu();
            st_relaxed(P.flags + blockIdx.x * R_FLAG_STRIDE, P.epoch0 + n + 1);
        }
        PROF(5);
    }
    if (prof) {
        atomicAdd(&g_resident_prof[6], (unsigned long long)P.n_pairs);
        atomicAdd(&g_resident_prof[7], (unsigned long long)(clock64() - t_launch));
    }
#undef PROF
}

struct ResChoice {
    bool ok;
    int nbx, nby, BX, BY, ZS;
    size_t smem;
};

constexpr int R_TABK = 256;      // class kinds: weight rows staged in shared memory (most frequent classes first)

static size_t resident_smem(int BX, int BY, int ZS, bool cls, int tab_k)
{
    const size_t tables = (size_t)(4 * (BY + 4) + 4 * BX) + 3 * (size_t)BX * BY + (size_t)(BX + 2) * (BY + 2);   // int2 entries
    const size_t codes = (size_t)(BX + 2) * (BY + 2) * (cls ? ZS : ZS / 4) * 2;
    return (size_t)(BX + 4) * (BY + 4) * ZS * 4 + codes + (cls ? (size_t)tab_k * 32 : 0) + tables * 8 + 32;
}

// Bricks: at most one per SM; smallest colour-A region ((BX + 2) x (BY + 2) rows per CTA), then the fewest frame rows.
static ResChoice choose_bricks(const taub_geom &g, int sms, bool cls, int tab_k)
{
    ResChoice best{};
    best.ok = false;
    const int ZS = ((g.Nz + 8 + 7) / 8) * 8;
    long best_cost = -1;
    for (int nbx = 1; nbx <= g.Nx / 2 && nbx * g.bs <= sms; ++nbx) {
        for (int nby = 1; nby <= g.Ny / 2 && (int64_t)nbx * nby * g.bs <= sms; ++nby) {
            const int BX = ceil_div(g.Nx, nbx), BY = ceil_div(g.Ny, nby);
            const size_t smem = resident_smem(BX, BY, ZS, cls, tab_k);
            if (smem > R_SMEM_MAX) continue;
            const long cost = (long)(BX + 2) * (BY + 2) * 64 + (BX + BY);
            if (best_cost < 0 || cost < best_cost) {
                best_cost = cost;
                best = ResChoice{true, nbx, nby, BX, BY, ZS, smem};
            }
        }
    }
    return best;
}

static int resident_env()
{
    static const int on = [] {
        const char *e = getenv("TAUB_RESIDENT");
        return (e && *e) ? atoi(e) : 1;
    }();
    return on;
}

}  // namespace taub

using namespace taub;

extern "C" {

unsigned long long taub_resident_timeouts(void)
{
    unsigned long long v = 0;
    if (cudaMemcpyFromSymbol(&v, g_resident_timeouts, sizeof(v)) != cudaSuccess) return ~0ULL;
    return v;
}

size_t taub_sync_ws_ints(void) { return (size_t)R_MAX_BRICKS * R_FLAG_STRIDE; }

int taub_resident_profile(unsigned long long out[8], int reset)
{
    TAUB_REQUIRE(out != nullptr, "taub_resident_profile: null pointer");
    TAUB_CUDA(cudaMemcpyFromSymbol(out, g_resident_prof, 8 * sizeof(unsigned long long)));
    if (reset) {
        const unsigned long long zero[8] = {0, 0, 0, 0, 0, 0, 0, 0};
        TAUB_CUDA(cudaMemcpyToSymbol(g_resident_prof, zero, sizeof(zero)));
    }
    return TAUB_OK;
}

static int device_sms(int *sms, int *coop)
{
    int dev = 0;
    TAUB_CUDA(cudaGetDevice(&dev));
    static int s_sms[64] = {0}, s_coop[64] = {0};
    if (!s_sms[dev & 63]) {
        TAUB_CUDA(cudaDeviceGetAttribute(&s_coop[dev & 63], cudaDevAttrCooperativeLaunch, dev));
        TAUB_CUDA(cudaDeviceGetAttribute(&s_sms[dev & 63], cudaDevAttrMultiProcessorCount, dev));
    }
    *sms = s_sms[dev & 63];
    *coop = s_coop[dev & 63];
    return TAUB_OK;
}

static bool res_cls(const taub_problem *p) { return p->kind == TAUB_MULTIPHASE_CLASS; }
static int res_tab_k(const taub_problem *p) { return res_cls(p) ? min(p->L, R_TABK) : 0; }

int taub_can_reside(const taub_problem *p)
{
    if (!p || (p->kind != TAUB_BINARY && p->kind != TAUB_MULTIPHASE_CLASS) || !p->codes || !p->field[0] || !p->field[1] ||
        !p->sync_ws)
        return 0;
    if (res_cls(p) && (!p->lut || p->L < 1)) return 0;
    if (!resident_env()) return 0;
    const taub_geom &g = p->g;
    if (g.i_offset != 0 || g.Nx != g.Nx_global) return 0;          // whole volumes only
    if (g.Nx < 2 || g.Ny < 2 || g.Nz < 8) return 0;
    if (g.periodic && ((g.Ny & 1) || (g.Nz & 1))) return 0;        // snapshot rule of the odd wrap: marching kernels
    if (p->peer_lo[0] || p->peer_hi[0] || p->peer_lo[1] || p->peer_hi[1]) return 0;
    int sms = 0, coop = 0;
    if (device_sms(&sms, &coop) != TAUB_OK || !coop) return 0;
    if ((int64_t)g.bs > sms) return 0;
    return choose_bricks(g, sms, res_cls(p), res_tab_k(p)).ok ? 1 : 0;
}

int taub_resident_pairs(taub_problem *p, int64_t iter, int n_pairs, void *stream)
{
    TAUB_REQUIRE(p && n_pairs >= 1, "taub_resident_pairs: bad arguments");
    if (taub_can_reside(p) != 1) {
        set_error("taub_resident_pairs: problem does not qualify for the shared-memory resident path");
        return TAUB_ERR_UNSUPPORTED;
    }
    const taub_geom &g = p->g;
    int sms = 0, coop = 0;
    if (int rc = device_sms(&sms, &coop)) return rc;
    const ResChoice c = choose_bricks(g, sms, res_cls(p), res_tab_k(p));
    const int bricks = g.bs * c.nbx * c.nby;
    TAUB_REQUIRE(bricks <= R_MAX_BRICKS, "taub_resident_pairs: more bricks than counters");
    ResParams P;
    P.g = g;
    P.buf[0] = p->field[p->cur];
    P.buf[1] = p->field[p->cur ^ 1];
    P.codes = p->codes;
    P.table = p->lut;
    P.n_classes = p->L;
    P.tab_k = res_tab_k(p);
    P.omega = p->omega;
    P.colour0 = (int)(iter & 1);
    P.n_pairs = n_pairs;
    P.nbx = c.nbx;
    P.nby = c.nby;
    P.BX = c.BX;
    P.BY = c.BY;
    P.ZS = c.ZS;
    P.flags = p->sync_ws;
    P.epoch0 = p->sync_epoch;
    P.stop = p->stop;
    static const int prof_on = [] {
        const char *e = getenv("TAUB_RESIDENT_PROF");
        return (e && *e) ? atoi(e) : 0;
    }();
    P.prof = prof_on;
    int dev = 0;
    TAUB_CUDA(cudaGetDevice(&dev));
    static const int nt = [] {
        const char *e = getenv("TAUB_RESIDENT_NT");
        return (e && atoi(e) == 1024) ? 1024 : 512;      // measured: 512 is 2-10 % faster at every size (32^3 .. 150^3)
    }();
    void *args[] = {(void *)&P};
#define TAUB_LAUNCH_RESIDENT(KIND_, NT_, BATCH_, SLOT_)                                                                    \
    do {                                                                                                                  \
        static bool attr_set[64] = {};                                                                                    \
        if (!attr_set[dev & 63]) {                                                                                        \
            TAUB_CUDA(cudaFuncSetAttribute(resident_kernel<KIND_, NT_, BATCH_>, cudaFuncAttributeMaxDynamicSharedMemorySize, \
                                           (int)R_SMEM_MAX));                                                             \
            attr_set[dev & 63] = true;                                                                                    \
        }                                                                                                                 \
        TAUB_CUDA(cudaLaunchCooperativeKernel((const void *)resident_kernel<KIND_, NT_, BATCH_>, dim3(bricks), dim3(NT_),  \
                                              args, c.smem, (cudaStream_t)stream));                                       \
    } while (0)
    if (res_cls(p)) {
        if (nt == 512) TAUB_LAUNCH_RESIDENT(TAUB_MULTIPHASE_CLASS, 512, 8, 0); else TAUB_LAUNCH_RESIDENT(TAUB_MULTIPHASE_CLASS, 1024, 4, 1);
    } else {
        if (nt == 512) TAUB_LAUNCH_RESIDENT(TAUB_BINARY, 512, 8, 2); else TAUB_LAUNCH_RESIDENT(TAUB_BINARY, 1024, 4, 3);
    }
#undef TAUB_LAUNCH_RESIDENT
    count_launch();
    p->sync_epoch += n_pairs;
    // the last pair stored the whole field into its buffer: buf[1] after an odd number of pairs, buf[0] otherwise
    if (n_pairs & 1) p->cur ^= 1;
    return TAUB_OK;
}

}  // extern "C"
