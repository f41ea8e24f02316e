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
//
// Round 2 ("instruction diet"): the step body exists twice -- a FAST instantiation for the steady state of the
// march (every plane-range flag true, no peer stores: straight-line code, predicated stores only) and the GENERIC
// one for the first / last few planes of a chunk; ring-slot addresses are rotated incrementally (no `% NB`), every
// shared-memory access is `thread offset + uniform slot offset + immediate`, the (n, 1/n) look-up is
// `(code >> k & 0x78) | table base` on a 128-byte aligned table, and threads without a column alias a real one
// instead of branching.  The stencil-class kinds keep the K most frequent weight rows in shared memory (the sweep's
// limiter was the L1 gather of those rows) with a warp-uniform vote per step for the rare others.
#include <stdlib.h>

#include <atomic>
#include <type_traits>

#include <cuda.h>   // CUtensorMap types only; the encoder is fetched through the runtime (no -lcuda)

#include "taub_common.cuh"

namespace taub {

constexpr int F_NT = 256;     // threads per CTA
constexpr int F_NRW = 4;      // rows per thread column
constexpr int F_NB = 6;       // field ring depth (planes in shared memory) of the binary kind
constexpr int F_NB_CLS = 4;   // ... of the class kinds (their slots carry 8 more bytes per float4 group)
constexpr int F_NBC_BIN = 4;  // code ring depth of the binary kind: codes are copied to registers when their plane
                              // is first read, so only the plane being read + 3 in flight need a slot
constexpr int F_TABK = 256;   // stencil-class rows kept in shared memory (most frequent classes first)
constexpr int REDO_CAP = 1022;  // chunks a redo list holds (2 header ints + REDO_CAP ids = 1024 ints); more: redo all
constexpr int REDO_LISTS = 4;   // lists per problem: passes on [0, Nx), [0, a), [b, Nx), [a, b) may be in flight together

struct FusedParams {
    taub_geom g;
    const float *src;
    float *dst;
    const uint16_t *codes;   // binary: one uint16 (four 4-bit counts) per group; class kind: one uint16 per voxel
    const float *table;      // class kind: [n_classes][8] weight rows
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
    int tab_k;         // class kind: rows of the weight table staged in shared memory (<= F_TABK)
    int write_solid;   // binary kind: 1 = also store float4 groups whose four voxels are all non-conductive
    int *redo;         // optional: redo list of this launch (see fused_redo_kernel)
    int force_redo;    // testing (TAUB_FORCE_REDO=1): list every chunk, i.e. the whole pass is redone with IEEE division
    int perm_R, perm_S;   // plane chunk of grid row y: (y % perm_R) * perm_S + y / perm_R (perm_R <= 1: y) -- the chunk rows
                          // that are in flight together lie perm_S chunks apart, and a chunk starts when its lower
                          // neighbour ends: the planes the two share are then still in L2
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

// Shared-memory accesses through 32-bit shared-window addresses: `thread register + uniform register + immediate`
// is one LDS / STS operand, so a ring-slot offset (uniform) and a compile-time row offset cost no instruction.
template <class T>
__device__ __forceinline__ T ld_sh(uint32_t a)
{
    return *reinterpret_cast<const T *>(__cvta_shared_to_generic((size_t)a));
}
template <class T>
__device__ __forceinline__ void st_sh(uint32_t a, const T &v)
{
    *reinterpret_cast<T *>(__cvta_shared_to_generic((size_t)a)) = v;
}
// 128-bit shared load that neither the compiler nor ptxas can split into scalar loads (a scalar load of one
// component of consecutive float4 groups is a 4-way bank conflict: same wavefronts, a quarter of the data).
template <int IMM>
__device__ __forceinline__ float4 lds128(uint32_t a)
{
    float4 v;
    asm volatile("ld.volatile.shared.v4.f32 {%0, %1, %2, %3}, [%4+%5];"
                 : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w)
                 : "r"(a), "n"(IMM));
    return v;
}

// (n, 1/n) of the 4-bit neighbour count at bit `SH` of `code`, from the 128-byte aligned static table at shared
// address `base`: the entry's byte offset n * 8 is OR-ed into the base (one shift, one LOP3, one LDS.64; no
// generic-pointer arithmetic).  Plain asm: the table is written once, before the first barrier of the kernel.
template <int SH>
__device__ __forceinline__ float2 div_pair_at(unsigned code, uint32_t base)
{
    const unsigned off = (SH >= 3) ? (code >> (SH - 3)) : (code << (3 - SH));
    float2 v;
    asm("ld.shared.v2.f32 {%0, %1}, [%2];" : "=f"(v.x), "=f"(v.y) : "r"((off & 0x78u) | base));
    return v;
}

// One row of a column, binary kind: "xz" rows update components x and z, "yw" rows y and w.  zs is the one z
// neighbour that lives in the adjacent group (.w of the left group / .x of the right group).  The row's 16-bit
// code word sits at bit SH16 (0 or 16) of `codes2`.
template <bool IS_XZ, int SH16, bool EXACT>
__device__ __forceinline__ void row_update(float4 &c, const float4 &xp, const float4 &xm, const float4 &up, const float4 &dn,
                                           float zs, unsigned codes2, uint32_t div_base, float omega, unsigned &umin)
{
    if (IS_XZ) {
        const float n0 = sor_fast<EXACT>(c.x, xp.x, xm.x, up.x, dn.x, c.y, zs, div_pair_at<SH16 + 0>(codes2, div_base), omega, umin);
        const float n1 = sor_fast<EXACT>(c.z, xp.z, xm.z, up.z, dn.z, c.w, c.y, div_pair_at<SH16 + 8>(codes2, div_base), omega, umin);
        c.x = n0;
        c.z = n1;
    } else {
        const float n0 = sor_fast<EXACT>(c.y, xp.y, xm.y, up.y, dn.y, c.z, c.x, div_pair_at<SH16 + 4>(codes2, div_base), omega, umin);
        const float n1 = sor_fast<EXACT>(c.w, xp.w, xm.w, up.w, dn.w, zs, c.z, div_pair_at<SH16 + 12>(codes2, div_base), omega, umin);
        c.y = n0;
        c.w = n1;
    }
}

// Weight row of a stencil class (two float4: {w_x+, w_x-, w_y+, w_y-}, {w_z+, w_z-, b, 1/b}): from the shared-memory
// copy of the tab_k most frequent rows or -- for the lanes whose class is a rare one -- through the read-only global
// path.  The compiler turns the two sides into complementary predicated loads: no divergence, and a predicated-off
// load costs an issue slot but no memory transaction.
struct ClassTab {
    const float4 *g;   // global table, two float4 per class
    uint32_t s;        // shared-window address of the staged copy of the first k rows
    unsigned k;        // rows staged
};

__device__ __forceinline__ void class_row(unsigned cls, const ClassTab &T, float4 &wa, float4 &wb)
{
    if (cls < T.k) {   // staged as two arrays of half rows: consecutive classes sit in consecutive 16-byte bank groups
        wa = ld_sh<float4>(T.s + cls * 16u);
        wb = ld_sh<float4>(T.s + (T.k + cls) * 16u);
    } else {
        const float4 *row = T.g + 2 * cls;
        wa = __ldg(row);
        wb = __ldg(row + 1);
    }
}

// Class kind: cls2 = the row's four uint16 class ids (x | y << 16, z | w << 16).
template <bool IS_XZ, bool EXACT>
__device__ __forceinline__ void row_update_class(float4 &c, const float4 &xp, const float4 &xm, const float4 &up,
                                                 const float4 &dn, float zs, uint2 cls2, const ClassTab &T, float omega,
                                                 unsigned &umin)
{
    const unsigned c0 = IS_XZ ? (cls2.x & 0xffffu) : (cls2.x >> 16), c1 = IS_XZ ? (cls2.y & 0xffffu) : (cls2.y >> 16);
    float4 wa0, wb0, wa1, wb1;
    class_row(c0, T, wa0, wb0);
    class_row(c1, T, wa1, wb1);
    if (IS_XZ) {
        const float n0 = sor_class_rows<EXACT>(c.x, xp.x, xm.x, up.x, dn.x, c.y, zs, wa0, wb0, omega, umin);
        const float n1 = sor_class_rows<EXACT>(c.z, xp.z, xm.z, up.z, dn.z, c.w, c.y, wa1, wb1, omega, umin);
        c.x = n0;
        c.z = n1;
    } else {
        const float n0 = sor_class_rows<EXACT>(c.y, xp.y, xm.y, up.y, dn.y, c.z, c.x, wa0, wb0, omega, umin);
        const float n1 = sor_class_rows<EXACT>(c.w, xp.w, xm.w, up.w, dn.w, zs, c.z, wa1, wb1, omega, umin);
        c.y = n0;
        c.w = n1;
    }
}

// Anisotropic kind: cls2 as above, the (b, 1/b) pair of a class from the static shared table.
template <bool IS_XZ, bool EXACT>
__device__ __forceinline__ void row_update_aniso(float4 &c, const float4 &xp, const float4 &xm, const float4 &up,
                                                 const float4 &dn, float zs, uint2 cls2, const float2 *s_div, float Ky, float Kz,
                                                 float omega, unsigned &umin)
{
    if (IS_XZ) {
        const float n0 = sor_aniso_fast<EXACT>(c.x, xp.x, xm.x, up.x, dn.x, c.y, zs, s_div[cls2.x & 0xffffu], Ky, Kz, omega, umin);
        const float n1 = sor_aniso_fast<EXACT>(c.z, xp.z, xm.z, up.z, dn.z, c.w, c.y, s_div[cls2.y & 0xffffu], Ky, Kz, omega, umin);
        c.x = n0;
        c.z = n1;
    } else {
        const float n0 = sor_aniso_fast<EXACT>(c.y, xp.y, xm.y, up.y, dn.y, c.z, c.x, s_div[cls2.x >> 16], Ky, Kz, omega, umin);
        const float n1 = sor_aniso_fast<EXACT>(c.w, xp.w, xm.w, up.w, dn.w, zs, c.z, s_div[cls2.y >> 16], Ky, Kz, omega, umin);
        c.y = n0;
        c.w = n1;
    }
}

template <int V>
using IC = std::integral_constant<int, V>;

// The march of one CTA over its chunk of planes (the whole kernel but its one-time set-up).  EXACT = false: the fast
// exact-reciprocal division; returns whether this thread met a non-zero value below 2^-100 where that division may be
// off by one subnormal ulp.  EXACT = true: IEEE division -- fused_redo_kernel re-runs the chunks of the CTAs that
// reported such a value (a kernel of its own, so the hot kernel's registers and instruction stream are untouched):
// the source buffer is read-only during a pass, so the re-run simply overwrites the first run's output.  The fused
// pass therefore equals IEEE division for every finite input.
template <int OGT, int PA0, int KIND, bool OP, bool EXACT>
__device__ __forceinline__ bool fused_march(const FusedParams &P, const CUtensorMap *tmap_p, const CUtensorMap *cmap_p,
                                            unsigned char *smem_raw, const float2 *s_div, int cta_x, int cta_y, int cta_z)
{
    const taub_geom &g = P.g;
    constexpr bool ANI = (KIND == TAUB_ANISOTROPIC);
    constexpr bool MPC = (KIND == TAUB_MULTIPHASE_CLASS);
    constexpr bool CLS = MPC || ANI;             // one uint16 id per voxel travels with the field
    constexpr int NRW = F_NRW;                   // rows per thread column
    constexpr int CPG = CLS ? 4 : 1;             // uint16 side-array elements per float4 group
    constexpr int NB = CLS ? F_NB_CLS : F_NB;    // field ring depth (planes of the tile resident in shared memory)
    constexpr int NBC = CLS ? NB : F_NBC_BIN;    // code ring depth
    constexpr int LGt = OGT + 2, LG = OGT + 3, LGc = OGT + 8;   // thread groups per row; box widths (field: odd pitch)
    constexpr uint32_t ROWB = LG * 16u, CROWB = LGc * CPG * 2u; // bytes per field row / id-code row of a ring slot
    const int LR = P.LR;
    const float4 *tab4 = reinterpret_cast<const float4 *>(P.table);   // class kind: two float4 per class
    const uint32_t slotB = (uint32_t)P.slot_f4 * 16u;   // bytes per field ring slot (multiple of 128)
    const uint32_t cslotB = (uint32_t)P.cslot_h * 2u;   // bytes per code ring slot (multiple of 128)
    unsigned char *planes = smem_raw;
    unsigned char *cplanes = smem_raw + (size_t)NB * slotB;
    uint64_t *mbar = reinterpret_cast<uint64_t *>(cplanes + (size_t)NBC * cslotB);
    float4 *s_tab = reinterpret_cast<float4 *>(mbar + 16);   // class kind: staged weight rows
    const uint32_t mbar_u32 = smem_u32(mbar);

    const int tid = threadIdx.x, lane = tid & 31;
    const int tk = cta_x % P.tiles_k, tj = cta_x / P.tiles_k;     // (cta_x, cta_y, cta_z): tile, plane chunk, image
    const int b = cta_z;
    const int c0 = P.i_lo + cta_y * P.chunk_len;
    const int c1 = min(c0 + P.chunk_len, P.i_hi);
    const int R0 = tj * P.OR_, G0 = tk * P.OG;   // storage row / group of loaded (0, 0)
    const int PG = g.pitch >> 2;
    const int total_rel = c1 - c0 + 4;           // planes c0-2 .. c1+1
    const int64_t ps = g.plane_stride;
    const ClassTab ctab{tab4, smem_u32(s_tab), (unsigned)P.tab_k};
    const float Ky = ANI ? P.table[2 * ANISO_CLASSES] : 0.0f, Kz = ANI ? P.table[2 * ANISO_CLASSES + 1] : 0.0f;
    const uint32_t div_base = smem_u32(s_div);
    unsigned umin = 0xffffffffu;   // guard word of every neighbour sum this thread divides

    // ---- TMA producer (thread 0): plane rel (local plane c0-2+rel) -> field slot rel % NB, code slot rel % NBC,
    //      one box each; the part of a box outside the tensor reads as 0.  The slot offsets and the plane
    //      coordinate of the next box are carried along instead of being recomputed from rel.
    const uint32_t planes_u32 = smem_u32(planes), cplanes_u32 = smem_u32(cplanes);
    const uint32_t tx_bytes = (uint32_t)LR * (ROWB + CROWB);
    uint32_t is_f = 0, is_c = 0, is_bar = mbar_u32;   // next issue: field / code slot byte offset, barrier address
    int is_pl = b * g.planes + (c0 - 2 + G);          // ... and tensor plane coordinate
    // with_codes = false: plane c0-2 (rel 0) is only ever an x- neighbour, nothing on it is updated, so its codes / ids
    // are never read -- and must not be loaded: the binary kind's code ring has fewer slots (4) than the prologue has
    // boxes in flight (5), and two boxes in flight into the same slot (rel 0 and rel 4) may land in either order
    auto issue = [&](const bool with_codes) {
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        mbar_expect_tx(is_bar, with_codes ? tx_bytes : (uint32_t)LR * ROWB);
        tma_load_3d(planes_u32 + is_f, tmap_p, 4 * G0, R0, is_pl, is_bar);
        if (with_codes) tma_load_3d(cplanes_u32 + is_c, cmap_p, G0 * CPG, R0, is_pl, is_bar);
        ++is_pl;
        is_f += slotB;
        is_bar += 8u;
        if (is_f == NB * slotB) {
            is_f = 0;
            is_bar = mbar_u32;
        }
        is_c += cslotB;
        if (is_c == NBC * cslotB) is_c = 0;
    };
    if (tid == 0)
        for (int rel = 0; rel < min(NB - 1, total_rel); ++rel) issue(rel > 0);

    // ---- this thread's column: rows lr0 .. lr0+NRW-1 of the tile, group gg.  Threads beyond the tile's columns
    //      (m >= NCT) shadow column 0 of their group: they load and compute like everybody else (no divergence in
    //      the step body) but never store anything.
    const int NCT = (LR - 2) / NRW;          // columns stacked in the tile
    const int m_raw = tid / LGt, gg = tid - m_raw * LGt;
    const bool own = m_raw < NCT;            // has cells in the shared-memory tile
    const int m = own ? m_raw : 0;
    const int lr0 = 1 + NRW * m;
    const int Ra = R0 + lr0, Gs = G0 + gg;
    const bool doit = own && (Gs < PG) && (Ra < g.rows);
    const bool colB = gg >= 1 && gg < LGt - 1 && Gs >= 1 && Gs < 1 + interior_groups(g.Nz);
    unsigned canB = 0;                       // bit r: row r of the column is an output row
#pragma unroll
    for (int r = 0; r < NRW; ++r)
        if (doit && colB && lr0 + r >= 2 && lr0 + r < LR - 2 && Ra + r >= G && Ra + r < G + g.Ny) canB |= 1u << r;
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
    // shared-window addresses of row 0 of the column in field slot 0 / id-code slot 0 (row r: + r * ROWB / CROWB)
    const uint32_t tbase = planes_u32 + (uint32_t)(lr0 * LG + gg) * 16u;
    const uint32_t cbase = cplanes_u32 + (uint32_t)((lr0 * LGc + gg) * CPG) * 2u;
    const bool lane_lo = (lane == 0), lane_hi = (lane == 31);
    // which of this thread's rows other threads read: the column's first and last row (their above / below) and,
    // at the two ends of a warp, every row (z_neighbour fall-back of the adjacent warp)
    const bool st_mid = own && (lane_lo || lane_hi);
    // colour B first writes plane c0 (at step 2)
    float *dst0 = P.dst + (int64_t)b * g.image_stride + 4 * Gs + (int64_t)(c0 + G) * ps + (int64_t)Ra * g.pitch;

    // ---- register ring: rg[r][k] holds plane (c0-3+k+4j) of row r; at step s = 4j+ss:
    //      a[p-2] = rg[.][ss], a[p-1] = rg[.][ss+1], raw[p] -> a[p] = rg[.][ss+2], raw[p+1] = rg[.][ss+3]
    float4 rg[NRW][4];
    unsigned cr[NRW / 2][4];   // binary: neighbour codes, two rows per word, same ring positions
    mbar_wait(mbar_u32, 0);
    mbar_wait(mbar_u32 + 8u * (1 % NB), 0);
#pragma unroll
    for (int r = 0; r < NRW; ++r) {
        rg[r][0] = make_float4(0.f, 0.f, 0.f, 0.f);
        rg[r][3] = make_float4(0.f, 0.f, 0.f, 0.f);
        rg[r][1] = ld_sh<float4>(tbase + r * ROWB);
        rg[r][2] = ld_sh<float4>(tbase + slotB + r * ROWB);
    }
#pragma unroll
    for (int q = 0; q < NRW / 2; ++q) {
#pragma unroll
        for (int k = 0; k < 4; ++k) cr[q][k] = 0;
        if (!CLS)
            cr[q][2] = (unsigned)ld_sh<uint16_t>(cbase + cslotB + 2 * q * CROWB) |
                       ((unsigned)ld_sh<uint16_t>(cbase + cslotB + (2 * q + 1) * CROWB) << 16);
    }

    const int n_steps = c1 - c0 + 2;

    // ---- ring state of step s (all warp-uniform): field slots of planes p-1, p, p+1 and the mbarrier of p+1;
    //      code / id slots of the same planes.  Rotated at the end of every step.
    uint32_t oM1 = 0, oP = slotB, oP1 = 2u * slotB;               // s % NB, (s+1) % NB, (s+2) % NB at s = 0
    uint32_t w_bar = mbar_u32 + 16u, w_par = 0;                      // barrier / parity of plane rel = s+2
    uint32_t kM1 = 0, kP = cslotB, kP1 = 2u * cslotB;                // code / id slots: s % NBC, ... (NBC >= 4)

    // One step of the march.  FAST: every plane-range flag is true and there are no peer stores (steady state).
    auto step = [&](auto fast_c, auto ss_c, const int s) {
        constexpr bool FAST = decltype(fast_c)::value != 0;
                constexpr int ss = decltype(ss_c)::value;
        constexpr int iM2 = ss & 3, iM1 = (ss + 1) & 3, iP = (ss + 2) & 3, iP1 = (ss + 3) & 3;
        const int p = c0 - 1 + s;   // plane receiving colour A; colour B goes to plane p-1
        __syncthreads();            // ring slot of plane p-2 is free; a[p-1] is visible in its slot
        if (tid == 0 && (FAST || s - 1 + NB < total_rel)) issue(true);
        mbar_wait(w_bar, w_par);
        const uint32_t aM1 = tbase + oM1, aP = tbase + oP, aP1 = tbase + oP1;   // this column in the three slots
        const bool doA = FAST || ((p >= P.a_lo) && (p < P.a_hi));
        const bool keepA = FAST || ((p >= c0) && (p < c1));   // a[p] is read by colour B of plane p next step
        const bool doB = FAST || (s >= 2);
        // one-sided halo exchange: output plane p-1 is one of the neighbour's ghost planes
        const bool send_lo = !FAST && P.peer_lo != nullptr && (p - 1) < G;
        const bool send_hi = !FAST && P.peer_hi != nullptr && (p - 1) >= g.Nx - G;
#pragma unroll
        for (int r = 0; r < NRW; ++r) rg[r][iP1] = ld_sh<float4>(aP1 + r * ROWB);
        if (!CLS) {
#pragma unroll
            for (int q = 0; q < NRW / 2; ++q)
                cr[q][iP1] = (unsigned)ld_sh<uint16_t>(cbase + kP1 + 2 * q * CROWB) |
                             ((unsigned)ld_sh<uint16_t>(cbase + kP1 + (2 * q + 1) * CROWB) << 16);
        }
        if (doA) {   // block-uniform
            // the z neighbour from the adjacent lane's registers (one crossbar pass instead of a 4-way conflicted
            // shared load); the two lanes at the warp ends read shared memory
            float zs[NRW];
#pragma unroll
            for (int r = 0; r < NRW; ++r) {
                if (((PA0 + ss + r) & 1) == 0) {
                    zs[r] = __shfl_up_sync(0xffffffffu, rg[r][iP].w, 1);
                    if (lane_lo) zs[r] = ld_sh<float>(aP + r * ROWB - 4);
                } else {
                    zs[r] = __shfl_down_sync(0xffffffffu, rg[r][iP].x, 1);
                    if (lane_hi) zs[r] = ld_sh<float>(aP + r * ROWB + 16);
                }
            }
            const float4 below = lds128<-(int)ROWB>(aP), above = lds128<(int)(NRW * ROWB)>(aP);
            uint2 ids[NRW];
            if (CLS) {
#pragma unroll
                for (int r = 0; r < NRW; ++r) ids[r] = ld_sh<uint2>(cbase + kP + r * CROWB);
            }
            float4 snap[NRW];
            if (OP) {
#pragma unroll
                for (int r = 0; r < NRW; ++r) snap[r] = rg[r][iP];
            }
#pragma unroll
            for (int r = 0; r < NRW; ++r) {
                // neighbours inside the column are registers; each row only reads the components
                // its neighbours leave unchanged in this step
                const float4 &dn = (r == 0) ? below : rg[r - 1][iP];
                const float4 &up = (r == NRW - 1) ? above : rg[r + 1][iP];
                if (((PA0 + ss + r) & 1) == 0) {
                    if (ANI)
                        row_update_aniso<true, EXACT>(rg[r][iP], rg[r][iP1], rg[r][iM1], up, dn, zs[r], ids[r], s_div, Ky, Kz, P.omega, umin);
                    else if (CLS)
                        row_update_class<true, EXACT>(rg[r][iP], rg[r][iP1], rg[r][iM1], up, dn, zs[r], ids[r], ctab, P.omega, umin);
                    else if (r & 1)
                        row_update<true, 16, EXACT>(rg[r][iP], rg[r][iP1], rg[r][iM1], up, dn, zs[r], cr[r >> 1][iP], div_base, P.omega, umin);
                    else
                        row_update<true, 0, EXACT>(rg[r][iP], rg[r][iP1], rg[r][iM1], up, dn, zs[r], cr[r >> 1][iP], div_base, P.omega, umin);
                } else {
                    if (ANI)
                        row_update_aniso<false, EXACT>(rg[r][iP], rg[r][iP1], rg[r][iM1], up, dn, zs[r], ids[r], s_div, Ky, Kz, P.omega, umin);
                    else if (CLS)
                        row_update_class<false, EXACT>(rg[r][iP], rg[r][iP1], rg[r][iM1], up, dn, zs[r], ids[r], ctab, P.omega, umin);
                    else if (r & 1)
                        row_update<false, 16, EXACT>(rg[r][iP], rg[r][iP1], rg[r][iM1], up, dn, zs[r], cr[r >> 1][iP], div_base, P.omega, umin);
                    else
                        row_update<false, 0, EXACT>(rg[r][iP], rg[r][iP1], rg[r][iM1], up, dn, zs[r], cr[r >> 1][iP], div_base, P.omega, umin);
                }
            }
            if (OP) {   // ghost cells of an odd periodic axis keep their snapshot
#pragma unroll
                for (int r = 0; r < NRW; ++r) {
                    if (keep_rows & (1u << r)) {
                        rg[r][iP] = snap[r];
                    } else {
                        if (keep_lo_w || keep_hi_w) rg[r][iP].w = snap[r].w;
                        if (keep_hi_y) rg[r][iP].y = snap[r].y;
                    }
                }
            }
            if (keepA) {
                // other threads read the column's first and last row (their above / below) and, at the two ends
                // of a warp, the neighbour lane's group (z_neighbour fall-back)
#pragma unroll
                for (int r = 0; r < NRW; ++r)
                    if ((r == 0 || r == NRW - 1) ? own : st_mid) st_sh<float4>(aP + r * ROWB, rg[r][iP]);
            }
        }
        if (doB) {   // block-uniform
            float zs[NRW];
#pragma unroll
            for (int r = 0; r < NRW; ++r) {
                if (((PA0 + ss + r) & 1) == 0) {
                    zs[r] = __shfl_up_sync(0xffffffffu, rg[r][iM1].w, 1);
                    if (lane_lo) zs[r] = ld_sh<float>(aM1 + r * ROWB - 4);
                } else {
                    zs[r] = __shfl_down_sync(0xffffffffu, rg[r][iM1].x, 1);
                    if (lane_hi) zs[r] = ld_sh<float>(aM1 + r * ROWB + 16);
                }
            }
            const float4 below = lds128<-(int)ROWB>(aM1), above = lds128<(int)(NRW * ROWB)>(aM1);
            uint2 ids[NRW];
            if (CLS) {
#pragma unroll
                for (int r = 0; r < NRW; ++r) ids[r] = ld_sh<uint2>(cbase + kM1 + r * CROWB);
            }
            float4 out[NRW];
#pragma unroll
            for (int r = 0; r < NRW; ++r) {
                out[r] = rg[r][iM1];
                const float4 &dn = (r == 0) ? below : rg[r - 1][iM1];
                const float4 &up = (r == NRW - 1) ? above : rg[r + 1][iM1];
                if (((PA0 + ss + r) & 1) == 0) {
                    if (ANI)
                        row_update_aniso<true, EXACT>(out[r], rg[r][iP], rg[r][iM2], up, dn, zs[r], ids[r], s_div, Ky, Kz, P.omega, umin);
                    else if (CLS)
                        row_update_class<true, EXACT>(out[r], rg[r][iP], rg[r][iM2], up, dn, zs[r], ids[r], ctab, P.omega, umin);
                    else if (r & 1)
                        row_update<true, 16, EXACT>(out[r], rg[r][iP], rg[r][iM2], up, dn, zs[r], cr[r >> 1][iM1], div_base, P.omega, umin);
                    else
                        row_update<true, 0, EXACT>(out[r], rg[r][iP], rg[r][iM2], up, dn, zs[r], cr[r >> 1][iM1], div_base, P.omega, umin);
                } else {
                    if (ANI)
                        row_update_aniso<false, EXACT>(out[r], rg[r][iP], rg[r][iM2], up, dn, zs[r], ids[r], s_div, Ky, Kz, P.omega, umin);
                    else if (CLS)
                        row_update_class<false, EXACT>(out[r], rg[r][iP], rg[r][iM2], up, dn, zs[r], ids[r], ctab, P.omega, umin);
                    else if (r & 1)
                        row_update<false, 16, EXACT>(out[r], rg[r][iP], rg[r][iM2], up, dn, zs[r], cr[r >> 1][iM1], div_base, P.omega, umin);
                    else
                        row_update<false, 0, EXACT>(out[r], rg[r][iP], rg[r][iM2], up, dn, zs[r], cr[r >> 1][iM1], div_base, P.omega, umin);
                }
            }
#pragma unroll
            for (int r = 0; r < NRW; ++r) {
                // binary kind: a float4 group whose four voxels are all non-conductive (code word 0) holds zeros in both
                // ping-pong buffers since the state build and for ever after -- there is nothing to write
                const bool live = CLS || P.write_solid || ((cr[r >> 1][iM1] >> (16 * (r & 1))) & 0xffffu) != 0u;
                if ((canB & (1u << r)) && live) {
                    float *d = dst0 + (int64_t)r * g.pitch;
                    *reinterpret_cast<float4 *>(d) = out[r];
                    if (send_lo) *reinterpret_cast<float4 *>(P.peer_lo + (d - P.dst) + (int64_t)g.Nx * ps) = out[r];
                    if (send_hi) *reinterpret_cast<float4 *>(P.peer_hi + (d - P.dst) - (int64_t)g.Nx * ps) = out[r];
                }
            }
            dst0 += ps;
        }
        // rotate the ring: plane p becomes p-1, ...; the slot after p+1's is the next to be waited for
        oM1 = oP;
        oP = oP1;
        oP1 += slotB;
        w_bar += 8u;
        if (oP1 == NB * slotB) {
            oP1 = 0;
            w_bar = mbar_u32;
            w_par ^= 1u;
        }
        kM1 = kP;
        kP = kP1;
        kP1 += cslotB;
        if (kP1 == NBC * cslotB) kP1 = 0;
    };

    // steps s4 .. s4+3 form a group (the register ring and the row parities have period 4): groups that lie wholly
    // in the steady state of the march -- colour A and B both active, plane kept, a box left to issue, no peer
    // stores -- run the FAST body
    const bool peers = (P.peer_lo != nullptr) || (P.peer_hi != nullptr);
    for (int s4 = 0; s4 < n_steps; s4 += 4) {
        // fast: s4 >= 2 (doB); p = c0-1+s in [max(a_lo, c0), min(a_hi, c1)) for s4..s4+3; s+NB-1 < total_rel
        const int p_first = c0 - 1 + s4, p_last = p_first + 3;
        bool fast = s4 >= 2 && p_first >= c0 && p_first >= P.a_lo && p_last < c1 && p_last < P.a_hi && s4 + 3 + NB - 1 < total_rel;
        if (peers && (p_first - 1 < G || p_last - 1 >= g.Nx - G)) fast = false;
        if (fast) {
            step(IC<1>{}, IC<0>{}, s4);
            step(IC<1>{}, IC<1>{}, s4 + 1);
            step(IC<1>{}, IC<2>{}, s4 + 2);
            step(IC<1>{}, IC<3>{}, s4 + 3);
        } else {
            step(IC<0>{}, IC<0>{}, s4);
            if (s4 + 1 < n_steps) step(IC<0>{}, IC<1>{}, s4 + 1);
            if (s4 + 2 < n_steps) step(IC<0>{}, IC<2>{}, s4 + 2);
            if (s4 + 3 < n_steps) step(IC<0>{}, IC<3>{}, s4 + 3);
        }
    }
    return !EXACT && doit && umin < GUARD_T;
}

// Thread work item = a COLUMN of NRW vertically adjacent rows x one float4 group.  In every step the
// rows of a column alternate between "xz" and "yw" rows and swap roles each step; the column's internal
// y-neighbours stay in registers, only the rows above and below it come from shared memory.  PA0 = parity
// of the column's first row at step 0 (uniform over the whole grid, chosen by the host), so every step
// body is branch-free.
//
// OP ("odd periodic"): a periodic extent Ny or Nz is odd, so the wrap joins two voxels of the SAME colour and a
// ghost cell is the image of a voxel whose colour differs from the ghost's own index parity.  The reference reads
// ghost SNAPSHOTS taken before each iteration (taufactor.py:501-505); with the images loaded once per pass that is
// reproduced exactly by leaving the ghost ring of the odd axis out of the colour-A step: where the imaged voxel has
// colour B the snapshot before iteration t+1 equals the loaded value, and where it has colour A no colour-B voxel
// reads it (rule pinned on the CPU by tests/test_fused_odd_periodic_cpu.py, on the GPU by the odd periodic goldens).
template <int OGT, int PA0, int KIND, bool OP = false>   // OGT: output groups per tile row (compile-time tile width);
__global__ void __launch_bounds__(F_NT, 2)              // KIND: TAUB_BINARY (4-bit codes) or a class kind (ids)
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
    constexpr bool ANI = (KIND == TAUB_ANISOTROPIC);
    constexpr bool MPC = (KIND == TAUB_MULTIPHASE_CLASS);
    constexpr bool CLS = MPC || ANI;
    constexpr int NB = CLS ? F_NB_CLS : F_NB, NBC = CLS ? NB : F_NBC_BIN;
    uint64_t *mbar = reinterpret_cast<uint64_t *>(smem_raw + (size_t)NB * P.slot_f4 * 16u + (size_t)NBC * P.cslot_h * 2u);
    float4 *s_tab = reinterpret_cast<float4 *>(mbar + 16);   // class kind: staged weight rows
    __shared__ __align__(512) float2 s_div[ANISO_CLASSES];   // static: constant address; aligned for the OR look-up
    const int tid = threadIdx.x;

    if (!ANI && tid < ANISO_CLASSES) s_div[tid] = div_entry(tid);   // binary: (n, 1/n) of the neighbour count
    const uint32_t mbar_u32 = smem_u32(mbar);
    if (tid == 0) {
        for (int n = 0; n < NB; ++n) mbar_init(mbar_u32 + 8u * n, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    // everything above touches only parameters and shared memory; from here on the kernel reads global memory
    // (stop flag, tables, source field): wait for the previous grid of the stream to complete and flush
    // (returns at once when this grid was not launched as a programmatic dependent)
    pdl_wait();
    if (P.stop && *P.stop) return;
    // anisotropic: (b, 1/b) of the prefactor classes; class kind: the first tab_k weight rows.  First used after
    // the step loop's first __syncthreads
    if (ANI && tid < ANISO_CLASSES) s_div[tid] = reinterpret_cast<const float2 *>(P.table)[tid];
    if (MPC) {
        const float4 *tab4 = reinterpret_cast<const float4 *>(P.table);
        for (int t = tid; t < 2 * P.tab_k; t += F_NT) s_tab[(t & 1) * P.tab_k + (t >> 1)] = __ldg(tab4 + t);
    }
    // (only the binary kind takes the elastic chunk model and with it the strided numbering: the other kinds, which have
    //  no register to spare, are compiled without it -- the class kind lost 4 % to these two lines)
    int cy = (int)blockIdx.y;
    if (KIND == TAUB_BINARY) {
        if (P.perm_R > 1) cy = ((int)blockIdx.y % P.perm_R) * P.perm_S + (int)blockIdx.y / P.perm_R;
        if (P.i_lo + cy * P.chunk_len >= P.i_hi) return;   // (a grid row the permutation pads the chunks with)
    }
    const bool tiny = fused_march<OGT, PA0, KIND, OP, false>(P, &tmap, &cmap, smem_raw, s_div, blockIdx.x, cy, blockIdx.z);
    // a value below 2^-100 went through the fast division: put this chunk on the list fused_redo_kernel works off
    if (__syncthreads_or(tiny || P.force_redo) && tid == 0) {
        if (P.redo) {
            const int k = atomicAdd(P.redo, 1);
            if (k < REDO_CAP) P.redo[2 + k] = (int)((blockIdx.z * gridDim.y + cy) * gridDim.x + blockIdx.x);
        }
        atomicAdd(&g_inexact_events, 1ULL);
    }
}

// Re-runs, with IEEE division, the chunks that the pass before it put on the redo list (P.redo: [0] = chunks listed,
// [1] = CTAs of this kernel that are done, [2..] = linear CTA ids of the pass; more than REDO_CAP entries = redo
// every chunk).  Launched behind every fused pass; all but never has anything to do (one load per CTA).
template <int OGT, int PA0, int KIND, bool OP = false>
__global__ void __launch_bounds__(F_NT, 1)
fused_redo_kernel(const FusedParams P, const __grid_constant__ CUtensorMap tmap, const __grid_constant__ CUtensorMap cmap,
                  int grid_x, int grid_y, int grid_z)
{
    pdl_trigger();
    extern __shared__ unsigned char smem_dyn[];
    unsigned char *smem_raw = smem_dyn + ((128u - (smem_u32(smem_dyn) & 127u)) & 127u);
    constexpr bool ANI = (KIND == TAUB_ANISOTROPIC);
    constexpr bool MPC = (KIND == TAUB_MULTIPHASE_CLASS);
    constexpr bool CLS = MPC || ANI;
    constexpr int NB = CLS ? F_NB_CLS : F_NB, NBC = CLS ? NB : F_NBC_BIN;
    uint64_t *mbar = reinterpret_cast<uint64_t *>(smem_raw + (size_t)NB * P.slot_f4 * 16u + (size_t)NBC * P.cslot_h * 2u);
    float4 *s_tab = reinterpret_cast<float4 *>(mbar + 16);
    __shared__ __align__(512) float2 s_div[ANISO_CLASSES];
    const int tid = threadIdx.x;
    pdl_wait();
    const int listed = *reinterpret_cast<volatile int *>(P.redo);
    if (listed == 0) return;                      // the common case
    if (!ANI && tid < ANISO_CLASSES) s_div[tid] = div_entry(tid);
    if (ANI && tid < ANISO_CLASSES) s_div[tid] = reinterpret_cast<const float2 *>(P.table)[tid];
    if (MPC) {
        const float4 *tab4 = reinterpret_cast<const float4 *>(P.table);
        for (int t = tid; t < 2 * P.tab_k; t += F_NT) s_tab[(t & 1) * P.tab_k + (t >> 1)] = __ldg(tab4 + t);
    }
    const uint32_t mbar_u32 = smem_u32(mbar);
    const int n = listed <= REDO_CAP ? listed : grid_x * grid_y * grid_z;
    bool armed = false;
    for (int k = blockIdx.x; k < n; k += gridDim.x) {
        const int id = listed <= REDO_CAP ? P.redo[2 + k] : k;
        if (P.i_lo + ((id / grid_x) % grid_y) * P.chunk_len >= P.i_hi) continue;   // (a padding row of the chunk permutation)
        __syncthreads();                          // every box of the previous chunk has been consumed
        if (tid == 0) {
            for (int m = 0; m < NB; ++m) {
                if (armed) asm volatile("mbarrier.inval.shared::cta.b64 [%0];" ::"r"(mbar_u32 + 8u * m) : "memory");
                mbar_init(mbar_u32 + 8u * m, 1);
            }
            asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        }
        armed = true;
        __syncthreads();
        fused_march<OGT, PA0, KIND, OP, true>(P, &tmap, &cmap, smem_raw, s_div, id % grid_x, (id / grid_x) % grid_y,
                                              id / (grid_x * grid_y));
    }
    // the last CTA to get here clears the list for the next pass
    __syncthreads();
    if (tid == 0) {
        __threadfence();
        if (atomicAdd(P.redo + 1, 1) == (int)gridDim.x - 1) {
            P.redo[1] = 0;
            __threadfence();
            P.redo[0] = 0;
        }
    }
}

static size_t fused_smem_bytes(int LR, int LG, int LGc, int cpg, int nb, int nbc, int tab_k)
{
    const size_t slot = ((size_t)(LR * LG + 7) / 8) * 8 * 16;               // fp32 box, 128-byte multiple
    const size_t cslot = ((size_t)(LR * LGc * cpg * 2 + 127) / 128) * 128;  // uint16 box (codes / class ids)
    return nb * slot + nbc * cslot + 16 * 8 + (size_t)tab_k * 32 + 128;      // + mbarriers (16 slots reserved), staged
                                                                             // weight rows, alignment slack (the
                                                                             // division table is static shared memory)
}

struct TileChoice {
    int LR, LG, LGc, LGt, OR_, OG, tiles_j, tiles_k;
    double eff;
};

// Tile = NCT columns (of F_NRW rows) stacked in y x LG float4 groups (OG of them are outputs, one halo group each
// side, one more to make the shared-memory row pitch odd: warps that straddle two columns would otherwise
// bank-conflict); one thread per (column, group).  TMA wants every box to start on a 16-byte boundary and to be a
// multiple of 16 bytes wide; for the uint16 code box that means OG (the tile step) is a multiple of 8 groups and the
// code box is OG + 8 wide.  The width is a template parameter of the kernel (row offsets become immediates), compiled
// for OG = 8, 16 and 32 (OG = 32 is the best shape of every extent that is a multiple of 128; the narrow ones serve
// 2-D images and small volumes); the height follows from the shared memory two CTAs per SM leave.  Pick the shape that
// wastes the fewest threads; wider wins a tie (longer rows per TMA box).
// Ring depth per kind.  The class kind carries 8 more bytes per float4 group in every slot and is bound by
// its weight-row look-ups, not by HBM latency: a shallower ring (one plane of prefetch) buys
// ~1.6x larger tiles -> fewer halo re-loads and idle threads (measured: 6 -> 4 slots = +10 % at 384^3 / 512^3).
static int ring_depth(int cpg) { return cpg == 1 ? F_NB : F_NB_CLS; }
static int code_ring_depth(int cpg) { return cpg == 1 ? F_NBC_BIN : F_NB_CLS; }

static TileChoice choose_tile(const taub_geom &g, int cpg, int tab_k, bool narrow_only)
{
    const int nb = ring_depth(cpg), nbc = code_ring_depth(cpg);
    // shared memory per CTA with two CTAs per SM: (228 KB - 2 x 1 KB reserved) / 2, less the 512-byte static
    // division table
    const size_t budget = 115712 - 512;
    const int ng = interior_groups(g.Nz);
    TileChoice best{};
    best.eff = -1.0;
    static const int shapes[3] = {8, 16, 32};
    for (int si = 0; si < (narrow_only ? 1 : 3); ++si) {
        const int OG = shapes[si];
        const int LGt = OG + 2, LG = OG + 3, LGc = OG + 8;
        int NCT = F_NT / LGt;   // columns the CTA's threads can cover
        while (NCT >= 1 && fused_smem_bytes(F_NRW * NCT + 2, LG, LGc, cpg, nb, nbc, tab_k) > budget) --NCT;
        if (NCT < 1) continue;
        const int NR = F_NRW * NCT, OR_ = NR - 2, LR = NR + 2;
        const int tj = ceil_div(g.Ny, OR_), tk = ceil_div(ng, OG);
        const double eff = ((double)g.Ny * ng) / ((double)tj * tk * F_NT * F_NRW);
        if (eff > best.eff - 1e-12) best = TileChoice{LR, LG, LGc, LGt, OR_, OG, tj, tk, eff};
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
    int dev, pitch, rows, planes_total, box_w, box_h, elem;
    bool operator==(const MapKey &o) const
    {
        return base == o.base && dev == o.dev && pitch == o.pitch && rows == o.rows && planes_total == o.planes_total &&
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
thread_local unsigned g_launch_cluster_x = 0;

// 3-D view of one ping-pong buffer: (columns = pitch, rows, bs * planes), fp32, box = LG*4 x LR x 1.
static int make_field_map(CUtensorMap *map, const taub_geom &g, const float *base, int LR, int LG, int dev)
{
    const MapKey key{base, dev, g.pitch, g.rows, g.bs * g.planes, LG * 4, LR, 4};
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
static int make_code_map(CUtensorMap *map, const taub_geom &g, const uint16_t *base, int LR, int LGc, int cpg, int dev)
{
    const MapKey key{base, dev, g.pitch * cpg / 4, g.rows, g.bs * g.planes, LGc * cpg, LR, 2};
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

// Plane chunks per tile column.  A CTA that marches `len` output planes is busy for about len + 6 plane steps (two
// more colour-A planes, the ring prologue, its launch; len + 4 where the prologue hides behind the SM's other CTA).  Two regimes, both measured (profiles/r2_chunks.md):
//   * list    -- CTAs run at a pace of their own (the kernel is bound by its instruction / shared-memory rate: the
//                class kinds, volumes that sit in L2): the cost is the makespan of handing the CTAs out in grid order
//                to `capacity` resident slots; the last chunk of a column may be short, which is what lets 300 CTAs
//                beat 260 on 296 slots;
//   * elastic -- the pass is HBM-bound (binary kind on a field beyond L2): resident CTAs share the memory system, so
//                fewer of them run faster and what counts is the work per slot plus a CTA of tail -- and the tiles of
//                a plane drift apart while they march, so the halo rows / columns a tile shares with its neighbours
//                miss L2 the more often the longer the chunk: on random voxels a plane step of a 22-plane chunk costs
//                0.75 of one of an 86-plane chunk (1024^3, profiles/r2_chunks_1024.txt; blob structures, which store
//                half as much, gain less).  Cost per plane step ~ 1 + len / 200, against 4 steps of overhead per chunk
//                (the planes neighbouring chunks load twice are mostly still in L2): 24 .. 28 planes at every size
//                from 384^3 to 2048^3.
static double chunk_cost(int n_planes, int64_t tiles, int capacity, bool elastic, int cl, int ce)
{
    const double oh = elastic ? 4.0 : 6.0;                              // (elastic: the ring prologue overlaps the SM's
    const double d = cl + oh, d2 = (n_planes - (ce - 1) * cl) + oh;     //  other CTA); full chunks, last chunk
    if (elastic) return ((double)tiles * ((ce - 1) * d + d2)) / capacity * (1.0 + cl / 200.0) + 0.25 * d;
    const int64_t n1 = tiles * (ce - 1);                 // CTAs of full length come first in grid order
    const int64_t R = n1 / capacity, r = n1 % capacity;
    const double t_full = (double)(R + (r > 0)) * d;     // when the last full-length CTA ends
    int64_t f = capacity - r;                            // slots free at R * d for the short CTAs
    double t_last;
    if (tiles <= f) {
        t_last = R * d + d2;
    } else {
        const int64_t k = (int64_t)(d / d2);             // rounds of short CTAs on those slots before the rest free up
        if (r > 0 && tiles > k * f)
            t_last = (R + 1) * d + (double)ceil_div64(tiles - k * f, capacity) * d2;
        else
            t_last = R * d + (double)ceil_div64(tiles, f) * d2;
    }
    return t_full > t_last ? t_full : t_last;
}

static void choose_chunks(int n_planes, int64_t tiles, int capacity, bool elastic, int *chunk_len, int *chunks)
{
    double best = 1e30;
    *chunk_len = n_planes + (n_planes & 1);
    *chunks = 1;
    for (int c = 1; c <= 512 && c <= (n_planes + 1) / 2; ++c) {
        int cl = ceil_div(n_planes, c);
        cl += cl & 1;   // even: every CTA starts with the same row parity
        const int ce = ceil_div(n_planes, cl);
        const double cost = chunk_cost(n_planes, tiles, capacity, elastic, cl, ce);
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

// Tuning switches (read once): TAUB_TABK = stencil-class rows staged in shared memory (default 256, 0 = none);
// TAUB_WRITE_SOLID=1 = store all-solid float4 groups too (the plain behaviour, for A/B measurements).
static int env_int(const char *name, int dflt)
{
    const char *e = getenv(name);
    return (e && *e) ? atoi(e) : dflt;
}
static int class_rows_in_smem(const taub_problem *p)
{
    static const int tabk = max(0, min(env_int("TAUB_TABK", F_TABK), 2048));
    return p->kind == TAUB_MULTIPHASE_CLASS ? min(p->L, tabk) : 0;
}

// The launch plan of a fused pass over planes [i_lo, i_hi): tile shape, plane chunks, grid, chunk numbering, cluster
// width.  Pure host arithmetic (no CUDA call): taub_fused_sweep2 launches what this returns, taub_fused_plan reports it.
struct FusedPlan {
    TileChoice t;
    int cpg, tab_k;
    bool op, elastic;
    int chunk_len, chunks;
    int grid_x, grid_y;
    int perm_R, perm_S;
    unsigned cluster_x;
};

static FusedPlan make_plan(const taub_problem *p, int i_lo, int i_hi, int resident_ctas)
{
    const taub_geom &g = p->g;
    FusedPlan f;
    f.cpg = (p->kind == TAUB_MULTIPHASE_CLASS || p->kind == TAUB_ANISOTROPIC) ? 4 : 1;
    f.tab_k = class_rows_in_smem(p);
    f.op = odd_periodic(g);   // the odd-periodic variant of the kernel is compiled for the narrow tile only
    f.t = choose_tile(g, f.cpg, f.tab_k, f.op);
    const int n_planes = i_hi - i_lo;
    const int64_t tiles = (int64_t)f.t.tiles_j * f.t.tiles_k * g.bs;
    // HBM-bound passes (binary kind, a field that does not sit in the 126 MB of L2: 320^3 and up) take the elastic
    // chunk model; the anisotropic and class kinds measured faster with the list model at every size
    const int chunk_model = env_int("TAUB_CHUNK_MODEL", -1);         // measurements: 0 = list, 1 = elastic
    f.elastic = chunk_model >= 0 ? chunk_model == 1
                                 : (p->kind == TAUB_BINARY && (double)g.bs * n_planes * g.plane_stride * 4.0 > 100e6);
    choose_chunks(n_planes, tiles, resident_ctas, f.elastic, &f.chunk_len, &f.chunks);
    const int chunks_env = env_int("TAUB_FUSED_CHUNKS", 0);          // measurements: plane chunks per tile column
    if (chunks_env > 0 && chunks_env <= (n_planes + 1) / 2) {
        f.chunk_len = ceil_div(n_planes, chunks_env);
        f.chunk_len += f.chunk_len & 1;
        f.chunks = ceil_div(n_planes, f.chunk_len);
    }
    f.grid_x = f.t.tiles_j * f.t.tiles_k;
    f.grid_y = f.chunks;
    // Elastic passes whose tiles do not fill the device by themselves run several chunk rows at a time.  Grid rows are
    // handed out in order, so with the plain numbering rows y and y+1 start together and the planes they share are
    // loaded a whole CTA lifetime apart (40 % L2 hits).  Numbered with a stride, the rows in flight lie perm_S chunks
    // apart and chunk y+1 starts when chunk y ends.  TAUB_FUSED_PERM=0: plain numbering (measurements).
    f.perm_R = 1;
    f.perm_S = f.chunks;
    if (f.elastic && p->kind == TAUB_BINARY && env_int("TAUB_FUSED_PERM", 1)) {
        const int R = min(f.chunks, ceil_div(resident_ctas, f.grid_x));
        if (R > 1) {
            f.perm_R = R;
            f.perm_S = ceil_div(f.chunks, R);
            f.grid_y = R * f.perm_S;      // (rows whose chunk lies beyond the last one return at once)
        }
    }
    // z-neighbour tiles are launched as clusters of two: co-scheduled CTAs start their march together, so the halo columns
    // they share are requested at about the same time and hit L2 (no cluster barrier, no distributed shared memory: the
    // kernel is unchanged).  512^3: blobs +0.6 %, random voxels +3 %, 256^3 +2 %; clusters of four: random +5 %, blobs
    // -3 % (a cluster waits for four free slots in one GPC); the class kind loses 1.5 % at 512^3, hence binary only.
    // TAUB_FUSED_CLUSTER=k overrides (0 / 1: none).
    static const unsigned cluster_env = (unsigned)env_int("TAUB_FUSED_CLUSTER", 2);
    f.cluster_x = (p->kind == TAUB_BINARY && cluster_env > 1 && f.grid_x % cluster_env == 0) ? cluster_env : 0;
    return f;
}

extern "C" {

unsigned long long taub_inexact_events(void)
{
    unsigned long long v = 0;
    if (cudaMemcpyFromSymbol(&v, g_inexact_events, sizeof(v)) != cudaSuccess) return ~0ULL;
    return v;
}

size_t taub_redo_ws_ints(void) { return (size_t)REDO_LISTS * (REDO_CAP + 2); }

int taub_can_fuse(const taub_problem *p)
{
    if (!p || (p->kind != TAUB_BINARY && p->kind != TAUB_MULTIPHASE_CLASS && p->kind != TAUB_ANISOTROPIC) || !p->codes ||
        !p->field[0] || !p->field[1])
        return 0;
    if (p->kind != TAUB_BINARY && !p->lut) return 0;
    const taub_geom &g = p->g;
    if (g.bs > 65535) return 0;
    return choose_tile(g, (p->kind == TAUB_MULTIPHASE_CLASS || p->kind == TAUB_ANISOTROPIC) ? 4 : 1, class_rows_in_smem(p),
                       odd_periodic(g)).eff > 0.0
               ? 1
               : 0;
}

int taub_fused_plan(const taub_problem *p, int i_lo, int i_hi, int resident_ctas, int32_t out[12])
{
    TAUB_REQUIRE(p && out && resident_ctas >= 1, "taub_fused_plan: bad arguments");
    if (taub_can_fuse(p) != 1) {
        set_error("taub_fused_plan: problem does not qualify for the fused path");
        return TAUB_ERR_UNSUPPORTED;
    }
    TAUB_REQUIRE(i_lo >= 0 && i_hi <= p->g.Nx && i_lo < i_hi, "taub_fused_plan: planes [%d, %d) outside the slab", i_lo, i_hi);
    const FusedPlan f = make_plan(p, i_lo, i_hi, resident_ctas);
    const int32_t v[12] = {f.t.LR, f.t.OR_, f.t.OG, f.t.tiles_j, f.t.tiles_k, f.chunk_len, f.chunks, f.grid_y,
                           f.perm_R, f.perm_S, (int32_t)f.cluster_x, f.elastic ? 1 : 0};
    for (int k = 0; k < 12; ++k) out[k] = v[k];
    return TAUB_OK;
}

int taub_fused_sweep2(const taub_problem *p, int64_t iter, int i_lo, int i_hi, void *stream)
{
    if (taub_can_fuse(p) != 1) {
        set_error("taub_fused_sweep2: problem does not qualify for the fused path");
        return TAUB_ERR_UNSUPPORTED;
    }
    const taub_geom &g = p->g;
    TAUB_REQUIRE(i_lo >= 0 && i_hi <= g.Nx && i_lo < i_hi, "taub_fused_sweep2: planes [%d, %d) outside the slab", i_lo, i_hi);
    int dev_ord = 0;
    TAUB_CUDA(cudaGetDevice(&dev_ord));
    static int sm_count[64] = {0};   // per device
    if (!sm_count[dev_ord & 63])
        TAUB_CUDA(cudaDeviceGetAttribute(&sm_count[dev_ord & 63], cudaDevAttrMultiProcessorCount, dev_ord));
    const FusedPlan plan = make_plan(p, i_lo, i_hi, 2 * sm_count[dev_ord & 63]);
    const TileChoice &t = plan.t;
    const int cpg = plan.cpg, tab_k = plan.tab_k;
    const bool op = plan.op;
    FusedParams P;
    P.g = g;
    P.src = p->field[p->cur];
    P.dst = p->field[p->cur ^ 1];
    P.codes = p->codes;
    P.table = p->lut;
    P.n_classes = p->L;
    P.tab_k = tab_k;
    static const int write_solid = env_int("TAUB_WRITE_SOLID", 0);
    P.write_solid = write_solid;
    P.omega = p->omega;
    P.stop = p->stop;
    // redo list of this launch: passes on the whole slab, its lower / upper boundary planes and its interior may be
    // in flight together (slab solver: boundary planes on a side stream), each kind on one stream at a time
    static const int exact_redo = env_int("TAUB_EXACT_REDO", 1);
    P.redo = (p->redo_ws && exact_redo)
                 ? p->redo_ws + (REDO_CAP + 2) * ((i_lo == 0 ? 0 : 2) + (i_hi == g.Nx ? 0 : 1))
                 : nullptr;
    {
        const char *e = getenv("TAUB_FORCE_REDO");      // read per launch: the tests switch it on and off
        P.force_redo = (e && *e && P.redo) ? atoi(e) : 0;
    }
    P.peer_lo = p->peer_lo[p->cur ^ 1];
    P.peer_hi = p->peer_hi[p->cur ^ 1];
    P.colourA = (int)(iter & 1);
    P.i_lo = i_lo;
    P.i_hi = i_hi;
    P.a_lo = max(i_lo - 1, -g.i_offset);
    P.a_hi = min(i_hi + 1, g.Nx_global - g.i_offset);
    P.LR = t.LR; P.LG = t.LG; P.LGc = t.LGc; P.LGt = t.LGt; P.OR_ = t.OR_; P.OG = t.OG;   // OR_ even, OG % 8 == 0
    P.tiles_k = t.tiles_k;
    P.chunk_len = plan.chunk_len;
    TAUB_REQUIRE(plan.grid_y <= 65535, "taub_fused_sweep2: too many plane chunks");
    P.slot_f4 = ((t.LR * t.LG + 7) / 8) * 8;
    P.cslot_h = ((t.LR * t.LGc * cpg * 2 + 127) / 128) * 64;
    const size_t smem = fused_smem_bytes(t.LR, t.LG, t.LGc, cpg, ring_depth(cpg), code_ring_depth(cpg), tab_k);
    dim3 grid(plan.grid_x, plan.grid_y, g.bs);
    P.perm_R = plan.perm_R;
    P.perm_S = plan.perm_S;
    const unsigned cluster_x = plan.cluster_x;
    cudaStream_t s = (cudaStream_t)stream;
    CUtensorMap tmap, cmap;
    if (int rc = make_field_map(&tmap, g, P.src, t.LR, t.LG, dev_ord)) return rc;
    if (int rc = make_code_map(&cmap, g, P.codes, t.LR, t.LGc, cpg, dev_ord)) return rc;
    // parity of loaded row 1 (row a of every pair) at step 0, i.e. at plane c0-1: 0 -> x,z active.
    // Tile row offsets (multiples of the even OR_) and chunk starts (multiples of the even
    // chunk_len) do not change it, so it is one number for the whole grid.
    const int pa0 = (1 - G + g.i_offset + P.colourA + (i_lo - 1)) & 1;   // row lr = 1 is row 0 of a column
    // the opt-in shared-memory limit of a kernel is per device and only ever raised: set it on every launch that
    // needs more than the largest value requested so far (a relaxed atomic per instantiation and device)
#define TAUB_LAUNCH_FUSED(OG_, PA_, KIND_, OP_)                                                                   \
    do {                                                                                                          \
        static std::atomic<size_t> smem_set[64];                                                                  \
        if (smem > smem_set[dev_ord & 63].load(std::memory_order_relaxed)) {                                      \
            TAUB_CUDA(cudaFuncSetAttribute(fused_sweep2_kernel<OG_, PA_, KIND_, OP_>,                             \
                                           cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));              \
            smem_set[dev_ord & 63].store(smem, std::memory_order_relaxed);                                        \
        }                                                                                                         \
        g_launch_cluster_x = cluster_x;                                                                           \
        cudaError_t le_ = launch_maybe_pdl(fused_sweep2_kernel<OG_, PA_, KIND_, OP_>, grid, dim3(F_NT), smem, s, P, tmap,  \
                                           cmap);                                                                 \
        g_launch_cluster_x = 0;                                                                                   \
        if (le_ != cudaSuccess && cluster_x > 1) {   /* a device / partition that cannot place the cluster */       \
            (void)cudaGetLastError();                                                                             \
            le_ = launch_maybe_pdl(fused_sweep2_kernel<OG_, PA_, KIND_, OP_>, grid, dim3(F_NT), smem, s, P, tmap, cmap); \
        }                                                                                                         \
        TAUB_CUDA(le_);                                                                                           \
        if (P.redo) {                                                                                             \
            static std::atomic<size_t> smem_set_r[64];                                                            \
            if (smem > smem_set_r[dev_ord & 63].load(std::memory_order_relaxed)) {                                \
                TAUB_CUDA(cudaFuncSetAttribute(fused_redo_kernel<OG_, PA_, KIND_, OP_>,                           \
                                               cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));          \
                smem_set_r[dev_ord & 63].store(smem, std::memory_order_relaxed);                                  \
            }                                                                                                     \
            /* an ordinary launch: as a programmatic dependent its CTAs (one per SM, a whole tile of shared memory  \
               each) would sit on the SMs while the pass is still running and take a slot from its later CTAs */   \
            fused_redo_kernel<OG_, PA_, KIND_, OP_><<<dim3(sm_count[dev_ord & 63]), dim3(F_NT), smem, s>>>(         \
                P, tmap, cmap, (int)grid.x, (int)grid.y, (int)grid.z);                                            \
            count_launch();                                                                                       \
        }                                                                                                         \
    } while (0)
#define TAUB_LAUNCH_FUSED_PA(OG_, KIND_, OP_)                                                                     \
    do {                                                                                                          \
        if (pa0 == 0) TAUB_LAUNCH_FUSED(OG_, 0, KIND_, OP_); else TAUB_LAUNCH_FUSED(OG_, 1, KIND_, OP_);          \
    } while (0)
#define TAUB_LAUNCH_FUSED_OG(KIND_)                                                                               \
    do {                                                                                                          \
        if (t.OG == 8) TAUB_LAUNCH_FUSED_PA(8, KIND_, false);                                                     \
        else if (t.OG == 16) TAUB_LAUNCH_FUSED_PA(16, KIND_, false);                                              \
        else TAUB_LAUNCH_FUSED_PA(32, KIND_, false);                                                              \
    } while (0)
    if (p->kind == TAUB_MULTIPHASE_CLASS) {
        if (op) TAUB_LAUNCH_FUSED_PA(8, TAUB_MULTIPHASE_CLASS, true); else TAUB_LAUNCH_FUSED_OG(TAUB_MULTIPHASE_CLASS);
    } else if (p->kind == TAUB_ANISOTROPIC) {
        TAUB_LAUNCH_FUSED_OG(TAUB_ANISOTROPIC);      // no periodic variant of this solver
    } else {
        if (op) TAUB_LAUNCH_FUSED_PA(8, TAUB_BINARY, true); else TAUB_LAUNCH_FUSED_OG(TAUB_BINARY);
    }
#undef TAUB_LAUNCH_FUSED_OG
#undef TAUB_LAUNCH_FUSED_PA
#undef TAUB_LAUNCH_FUSED
    TAUB_CUDA(cudaGetLastError());
    count_launch();
    return TAUB_OK;
}

}  // extern "C"
