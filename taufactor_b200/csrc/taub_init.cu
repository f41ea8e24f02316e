// taub_init.cu -- construction kernels: slab storage, neighbour codes / phase indices, counts.
//
// Replaces the state build of SORSolver.__init__ (taufactor.py:40-59): initial field
// (init_field :282-291), conductive-neighbour prefactor (:402-410, :493-499, :585-604, :626-650)
// and the per-slice volume-fraction numerators (:42).  One pass over the uint8 label image; no
// fp32 image copy, no meshgrid, no chequerboard tensors.
#include <stddef.h>

#include "taub_common.cuh"

namespace taub {

struct ImgView {
    const uint8_t *img;  // [bs][img_n][Ny][Nz], global planes [i0, i0 + n)
    int i0, n, Ny, Nz, Nx_global, periodic;
    int64_t image_stride;  // n * Ny * Nz
};

// Raw label of voxel (b, i, j, k) in GLOBAL x index; j/k must already be inside [0,Ny)x[0,Nz).
__device__ __forceinline__ int raw_label(const ImgView &v, int b, int i, int j, int k)
{
    return v.img[(int64_t)b * v.image_stride + ((int64_t)(i - v.i0) * v.Ny + j) * v.Nz + k];
}

// Weight a neighbour contributes to the binary neighbour count (taufactor.py:404-405, :494-495):
// the two Dirichlet ghost planes count 2, y/z outside counts 0 or wraps, inside = the mask.
__device__ __forceinline__ int nn_weight(const ImgView &v, int b, int i, int j, int k)
{
    if (i < 0 || i >= v.Nx_global) return 2;
    if (j < 0 || j >= v.Ny || k < 0 || k >= v.Nz) {
        if (!v.periodic) return 0;
        j = wrap(j, v.Ny);
        k = wrap(k, v.Nz);
    }
    return raw_label(v, b, i, j, k) == 1;
}

// One thread per float4 group of the slab storage.
__global__ void __launch_bounds__(256)
init_binary_kernel(taub_geom g, ImgView v, const float *__restrict__ vec, float *__restrict__ f0,
                   float *__restrict__ f1, uint16_t *__restrict__ codes)
{
    const int ngroups = g.pitch >> 2;
    const int64_t total = (int64_t)g.bs * g.planes * g.rows * ngroups;
    for (int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; t < total;
         t += (int64_t)gridDim.x * blockDim.x) {
        const int grp = (int)(t % ngroups);
        int64_t r = t / ngroups;
        const int jr = (int)(r % g.rows);
        r /= g.rows;
        const int ip = (int)(r % g.planes);
        const int b = (int)(r / g.planes);
        const int i = ip - G + g.i_offset;  // global x
        const int j = jr - G;
        float val[4];
        unsigned code = 0;
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            int k = grp * 4 + q - COL0;
            int jj = j;
            bool present = (k >= -G && k < g.Nz + G);  // interior or ghost frame (not row padding)
            const bool inside = (jj >= 0 && jj < g.Ny && k >= 0 && k < g.Nz);
            if (present && !inside) {
                if (g.periodic) {
                    jj = wrap(jj, g.Ny);
                    k = wrap(k, g.Nz);
                } else {
                    present = false;
                }
            }
            float x = 0.0f;
            unsigned c = 0;
            if (present) {
                if (i < 0) {
                    x = -1.0f;  // 2 * top_bc, taufactor.py:279, :291
                } else if (i >= g.Nx_global) {
                    x = 1.0f;   // 2 * bot_bc
                } else {
                    const int m = raw_label(v, b, i, jj, k) == 1;
                    x = __fmul_rn(m ? 1.0f : 0.0f, vec[i]);  // mask * linspace (keeps -0.0)
                    if (m) {
                        c = nn_weight(v, b, i + 1, jj, k) + nn_weight(v, b, i - 1, jj, k) +
                            nn_weight(v, b, i, jj + 1, k) + nn_weight(v, b, i, jj - 1, k) +
                            nn_weight(v, b, i, jj, k + 1) + nn_weight(v, b, i, jj, k - 1);
                        if (c == 0) c = 9;   // conductive but isolated: factor = inf, yet part of the mask
                    }
                }
            }
            val[q] = x;
            code |= c << (4 * q);
        }
        const float4 out = make_float4(val[0], val[1], val[2], val[3]);
        reinterpret_cast<float4 *>(f0)[t] = out;
        reinterpret_cast<float4 *>(f1)[t] = out;
        codes[t] = (uint16_t)code;
    }
}

// Multi-phase: dense phase index per storage voxel.  Ghost frame: wrapped (periodic) or the
// isolating pseudo-phase L; the two Dirichlet ghost planes copy the adjacent plane
// (taufactor.py:590-592, :631-638).  One thread per storage voxel quad.
__global__ void __launch_bounds__(256)
init_multi_kernel(taub_geom g, ImgView v, int L, const uint8_t *__restrict__ map256,
                  const float *__restrict__ cond, const float *__restrict__ vec,
                  float *__restrict__ f0, float *__restrict__ f1, uint8_t *__restrict__ labels)
{
    const int ngroups = g.pitch >> 2;
    const int64_t total = (int64_t)g.bs * g.planes * g.rows * ngroups;
    for (int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; t < total;
         t += (int64_t)gridDim.x * blockDim.x) {
        const int grp = (int)(t % ngroups);
        int64_t r = t / ngroups;
        const int jr = (int)(r % g.rows);
        r /= g.rows;
        const int ip = (int)(r % g.planes);
        const int b = (int)(r / g.planes);
        const int i = ip - G + g.i_offset;
        const int j = jr - G;
        float val[4];
        unsigned packed = 0;
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            int k = grp * 4 + q - COL0;
            int jj = j;
            bool present = (k >= -G && k < g.Nz + G);
            const bool inside = (jj >= 0 && jj < g.Ny && k >= 0 && k < g.Nz);
            if (present && !inside) {
                if (g.periodic) {
                    jj = wrap(jj, g.Ny);
                    k = wrap(k, g.Nz);
                } else {
                    present = false;
                }
            }
            float x = 0.0f;
            unsigned lab = (unsigned)L;
            if (present) {
                const int ic = min(max(i, 0), g.Nx_global - 1);  // Dirichlet ghosts copy the edge plane
                lab = map256[raw_label(v, b, ic, jj, k)];
                if (i < 0)
                    x = -1.0f;
                else if (i >= g.Nx_global)
                    x = 1.0f;
                else
                    x = __fmul_rn(cond[lab], vec[i]);
            }
            val[q] = x;
            packed |= lab << (8 * q);
        }
        const float4 out = make_float4(val[0], val[1], val[2], val[3]);
        reinterpret_cast<float4 *>(f0)[t] = out;
        reinterpret_cast<float4 *>(f1)[t] = out;
        reinterpret_cast<uint32_t *>(labels)[t] = packed;
    }
}

// counts[b][i] = voxels of local plane i whose raw label is selected; hist[b][256] optional.
__global__ void __launch_bounds__(256)
plane_counts_kernel(taub_geom g, ImgView v, const uint8_t *__restrict__ sel256,
                    unsigned long long *__restrict__ counts, unsigned long long *__restrict__ hist)
{
    __shared__ unsigned s_hist[256];
    __shared__ uint8_t s_sel[256];
    __shared__ unsigned s_cnt;
    const int b = blockIdx.y, il = blockIdx.x;
    const int i = il + g.i_offset;
    for (int t = threadIdx.x; t < 256; t += blockDim.x) {
        s_hist[t] = 0;
        s_sel[t] = sel256[t];
    }
    if (threadIdx.x == 0) s_cnt = 0;
    __syncthreads();
    const int64_t n = (int64_t)g.Ny * g.Nz;
    const uint8_t *plane = v.img + (int64_t)b * v.image_stride + (int64_t)(i - v.i0) * n;
    unsigned local = 0;
    for (int64_t t = threadIdx.x; t < n; t += blockDim.x) {
        const int lab = plane[t];
        local += s_sel[lab] != 0;
        if (hist) atomicAdd(&s_hist[lab], 1u);
    }
    for (int o = 16; o > 0; o >>= 1) local += __shfl_xor_sync(0xffffffffu, local, o);
    if ((threadIdx.x & 31) == 0) atomicAdd(&s_cnt, local);
    __syncthreads();
    if (threadIdx.x == 0) counts[(int64_t)b * g.Nx + il] = s_cnt;
    if (hist)
        for (int t = threadIdx.x; t < 256; t += blockDim.x)
            if (s_hist[t]) atomicAdd(&hist[(int64_t)b * 256 + t], (unsigned long long)s_hist[t]);
}

// One thread per interior voxel: pack the seven dense phase indices of its stencil.
__global__ void __launch_bounds__(256)
multiphase_keys_kernel(taub_geom g, const uint8_t *__restrict__ labels, int32_t *__restrict__ keys, int i_lo, int n_i)
{
    const int64_t total = (int64_t)g.bs * n_i * g.Ny * g.Nz;
    for (int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; t < total; t += (int64_t)gridDim.x * blockDim.x) {
        const int k = (int)(t % g.Nz);
        int64_t r = t / g.Nz;
        const int j = (int)(r % g.Ny);
        r /= g.Ny;
        const int i = i_lo + (int)(r % n_i);
        const int b = (int)(r / n_i);
        const int64_t o = (int64_t)b * g.image_stride + (int64_t)(i + G) * g.plane_stride + (int64_t)(j + G) * g.pitch + COL0 + k;
        const int ig = i + g.i_offset;
        int key = labels[o] | (labels[o - g.plane_stride] << 4) | (labels[o + g.plane_stride] << 8) |
                  (labels[o - g.pitch] << 12) | (labels[o + g.pitch] << 16) | (labels[o - 1] << 20) | (labels[o + 1] << 24);
        if (ig == 0) key |= 1 << 28;
        if (ig == g.Nx_global - 1) key |= 1 << 29;
        keys[t] = key;
    }
}

// ------------------------------------------------------------------------------------------------------
// Stencil classes of the multi-phase solvers, built on the device (no per-voxel key array, no sort of N^3 keys):
// a hash table of the distinct 30-bit stencil keys with their voxel counts (class_count_kernel), a host step on the
// <= 65534 distinct keys (ranking by frequency, weight rows in the reference's fp32 order), then one pass that
// writes the class id of every storage voxel (class_assign_kernel: interior, periodic frame, inert elsewhere).
// Workspace layout: int32 header[4] = {distinct keys, overflow flag, capacity, 0}; int32 keys[capacity] (-1 = empty);
// uint64 counts[capacity].
// ------------------------------------------------------------------------------------------------------
constexpr int CLS_LOG2 = 17, CLS_CAP = 1 << CLS_LOG2;
constexpr int CLS_MAX = 65534;      // class ids are uint16; one more id is the inert class

struct ClassWs {
    int n_distinct, overflow, capacity, pad;
    int keys[CLS_CAP];
    unsigned long long counts[CLS_CAP];
};

// The seven dense phase indices that determine the stencil of local voxel (i, j, k), + first / last global plane.
__device__ __forceinline__ int stencil_key(const taub_geom &g, const uint8_t *__restrict__ labels, int b, int i, int j, int k)
{
    const int64_t o = (int64_t)b * g.image_stride + (int64_t)(i + G) * g.plane_stride + (int64_t)(j + G) * g.pitch + COL0 + k;
    const int ig = i + g.i_offset;
    int key = labels[o] | (labels[o - g.plane_stride] << 4) | (labels[o + g.plane_stride] << 8) |
              (labels[o - g.pitch] << 12) | (labels[o + g.pitch] << 16) | (labels[o - 1] << 20) | (labels[o + 1] << 24);
    if (ig == 0) key |= 1 << 28;
    if (ig == g.Nx_global - 1) key |= 1 << 29;
    return key;
}

__device__ __forceinline__ unsigned class_hash(int key) { return ((unsigned)key * 2654435761u) >> (32 - CLS_LOG2); }

__global__ void __launch_bounds__(256)
class_count_kernel(taub_geom g, const uint8_t *__restrict__ labels, ClassWs *__restrict__ ws, int i_lo, int n_i)
{
    const int64_t total = (int64_t)g.bs * n_i * g.Ny * g.Nz;
    const int lane = threadIdx.x & 31;
    for (int64_t base = blockIdx.x * (int64_t)blockDim.x; base < total; base += (int64_t)gridDim.x * blockDim.x) {
        if (__any_sync(0xffffffffu, *(volatile int *)&ws->overflow)) return;   // warp-uniform: the vote below needs whole warps
        const int64_t t = base + threadIdx.x;
        int key = -2;                                    // lanes past the end share a key nobody inserts
        if (t < total) {
            const int k = (int)(t % g.Nz);
            int64_t r = t / g.Nz;
            const int j = (int)(r % g.Ny);
            r /= g.Ny;
            key = stencil_key(g, labels, (int)(r / n_i), i_lo + (int)(r % n_i), j, k);
        }
        // neighbouring voxels mostly share their stencil: one table update per distinct key of the warp
        const unsigned peers = __match_any_sync(0xffffffffu, key);
        if (key >= 0 && lane == __ffs(peers) - 1) {
            const unsigned long long n = (unsigned long long)__popc(peers);
            const unsigned h = class_hash(key);
            bool placed = false;
            for (int probe = 0; probe < CLS_CAP && !placed; ++probe) {
                const unsigned slot = (h + probe) & (CLS_CAP - 1);
                const int prev = atomicCAS(&ws->keys[slot], -1, key);
                if (prev == -1 && atomicAdd(&ws->n_distinct, 1) >= CLS_MAX) ws->overflow = 1;
                if (prev == -1 || prev == key) {
                    atomicAdd(&ws->counts[slot], n);
                    placed = true;
                }
            }
            if (!placed) ws->overflow = 1;
        }
    }
}

// One thread per float4 group of the slab storage: four class ids (one 8-byte store).
__global__ void __launch_bounds__(256)
class_assign_kernel(taub_geom g, const uint8_t *__restrict__ labels, const ClassWs *__restrict__ ws,
                    const uint16_t *__restrict__ slot_class, unsigned inert, uint16_t *__restrict__ classes, int i_lo, int i_hi)
{
    const int ngroups = g.pitch >> 2;
    const int64_t total = (int64_t)g.bs * g.planes * g.rows * ngroups;
    for (int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; t < total; t += (int64_t)gridDim.x * blockDim.x) {
        const int grp = (int)(t % ngroups);
        int64_t r = t / ngroups;
        const int jr = (int)(r % g.rows);
        r /= g.rows;
        const int i = (int)(r % g.planes) - G;
        const int b = (int)(r / g.planes);
        unsigned id[4];
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            int k = grp * 4 + q - COL0, j = jr - G;
            bool present = i >= i_lo && i < i_hi && k >= -G && k < g.Nz + G;      // interior or ghost frame
            if (present && !(j >= 0 && j < g.Ny && k >= 0 && k < g.Nz)) {
                if (g.periodic) {      // the fused kernel updates the first ghost ring like its periodic image
                    j = wrap(j, g.Ny);
                    k = wrap(k, g.Nz);
                } else {
                    present = false;
                }
            }
            id[q] = inert;
            if (present) {
                const int key = stencil_key(g, labels, b, i, j, k);
                unsigned slot = class_hash(key);
                for (int probe = 0; probe < CLS_CAP; ++probe, slot = (slot + 1) & (CLS_CAP - 1)) {
                    const int have = ws->keys[slot];
                    if (have == key) {
                        id[q] = slot_class[slot];
                        break;
                    }
                    if (have == -1) break;      // cannot happen for a key the count pass saw
                }
            }
        }
        reinterpret_cast<uint2 *>(classes)[t] = make_uint2(id[0] | (id[1] << 16), id[2] | (id[3] << 16));
    }
}

// ------------------------------------------------------------------------------------------------------
// Neighbour-count classes: the prefactor of the AnisotropicSolver (taufactor.py:459-471) and of the electrode solvers
// (cond_nn + k0 * reac_nn, taufactor.py:47-56 with electrode.py:48-64) depends only on a few small neighbour counts,
// so the state is one uint16 class id per storage voxel + a small table.  One pass over the label image builds the
// ids and the start field.
//   MODE 0 (anisotropic): id = (nx * 3 + ny) * 3 + nz, nx = conductive x neighbours with the Dirichlet planes counting
//     2 (0..4), ny / nz in 0..2 (no-flux faces); 63 = inert (non-conductive, outside).  Field as the binary solver.
//   MODE 1 (electrode): id = b * 112 + (cond_nn * 7 + reac_nn) * 2 + [x+ neighbour conducts]; cond_nn counts the
//     left Dirichlet plane twice, the right end is closed; y/z faces closed or periodic; inert = bs * 112.
//     Field: conductive voxels start at vec[i] (the ideal cosh profile), the left ghost plane holds 2 * left_bc = 2,
//     everything else 0.  reac_sums[b][i] += reactive-neighbour counts of the conductive voxels of plane i (a_x).
// ------------------------------------------------------------------------------------------------------
constexpr int ELECTRODE_IDS = 8 * 7 * 2;

template <int MODE>
__global__ void __launch_bounds__(256)
init_classes_kernel(taub_geom g, ImgView v, const float *__restrict__ vec, float *__restrict__ f0, float *__restrict__ f1,
                    uint16_t *__restrict__ ids, int cond_label, int reac_label, unsigned long long *__restrict__ reac_sums)
{
    const int ngroups = g.pitch >> 2;
    const int64_t total = (int64_t)g.bs * g.planes * g.rows * ngroups;
    const unsigned inert = MODE == 0 ? 63u : (unsigned)(g.bs * ELECTRODE_IDS);
    // label test of voxel (b, i, j, k) with the solver's boundary rules; `ghost_lo` = what the plane below x = 0 counts
    auto is_label = [&](int b, int i, int j, int k, int label, int ghost_lo, int ghost_hi) -> int {
        if (i < 0) return ghost_lo;
        if (i >= v.Nx_global) return ghost_hi;
        if (j < 0 || j >= v.Ny || k < 0 || k >= v.Nz) {
            if (!v.periodic) return 0;
            j = wrap(j, v.Ny);
            k = wrap(k, v.Nz);
        }
        return raw_label(v, b, i, j, k) == label;
    };
    for (int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; t < total; t += (int64_t)gridDim.x * blockDim.x) {
        const int grp = (int)(t % ngroups);
        int64_t r = t / ngroups;
        const int jr = (int)(r % g.rows);
        r /= g.rows;
        const int i = (int)(r % g.planes) - G + g.i_offset;   // global x
        const int b = (int)(r / g.planes);
        float val[4];
        unsigned id[4];
        unsigned long long reac_here = 0;
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            int k = grp * 4 + q - COL0, j = jr - G;
            bool present = (k >= -G && k < g.Nz + G);
            const bool inside = (j >= 0 && j < g.Ny && k >= 0 && k < g.Nz);
            if (present && !inside) {
                if (g.periodic) {
                    j = wrap(j, g.Ny);
                    k = wrap(k, g.Nz);
                } else {
                    present = false;
                }
            }
            float x = 0.0f;
            unsigned c = inert;
            if (present) {
                if (MODE == 0) {
                    if (i < 0) {
                        x = -1.0f;
                    } else if (i >= g.Nx_global) {
                        x = 1.0f;
                    } else {
                        const int m = raw_label(v, b, i, j, k) == 1;
                        x = __fmul_rn(m ? 1.0f : 0.0f, vec[i]);
                        if (m) {
                            const int nx = is_label(b, i + 1, j, k, 1, 2, 2) + is_label(b, i - 1, j, k, 1, 2, 2);
                            const int ny = is_label(b, i, j + 1, k, 1, 2, 2) + is_label(b, i, j - 1, k, 1, 2, 2);
                            const int nz = is_label(b, i, j, k + 1, 1, 2, 2) + is_label(b, i, j, k - 1, 1, 2, 2);
                            c = (unsigned)((nx * 3 + ny) * 3 + nz);
                        }
                    }
                } else {
                    if (i == -1) {
                        x = 2.0f;                                   // 2 * left_bc: the Dirichlet plane counts twice
                    } else if (i >= 0 && i < g.Nx_global && raw_label(v, b, i, j, k) == cond_label) {
                        x = vec[i];
                        const int xp = is_label(b, i + 1, j, k, cond_label, 2, 0);
                        const int cn = xp + is_label(b, i - 1, j, k, cond_label, 2, 0) + is_label(b, i, j + 1, k, cond_label, 2, 0) +
                                       is_label(b, i, j - 1, k, cond_label, 2, 0) + is_label(b, i, j, k + 1, cond_label, 2, 0) +
                                       is_label(b, i, j, k - 1, cond_label, 2, 0);
                        const int rn = is_label(b, i + 1, j, k, reac_label, 0, 0) + is_label(b, i - 1, j, k, reac_label, 0, 0) +
                                       is_label(b, i, j + 1, k, reac_label, 0, 0) + is_label(b, i, j - 1, k, reac_label, 0, 0) +
                                       is_label(b, i, j, k + 1, reac_label, 0, 0) + is_label(b, i, j, k - 1, reac_label, 0, 0);
                        c = (unsigned)(b * ELECTRODE_IDS + (cn * 7 + rn) * 2 + xp);
                        if (inside) reac_here += (unsigned long long)rn;
                    }
                }
            }
            val[q] = x;
            id[q] = c;
        }
        const float4 out = make_float4(val[0], val[1], val[2], val[3]);
        reinterpret_cast<float4 *>(f0)[t] = out;
        reinterpret_cast<float4 *>(f1)[t] = out;
        reinterpret_cast<uint2 *>(ids)[t] = make_uint2(id[0] | (id[1] << 16), id[2] | (id[3] << 16));
        if (MODE == 1 && reac_here) atomicAdd(&reac_sums[(int64_t)b * g.Nx_global + i], reac_here);
    }
}

static int check_img_cover(const taub_geom &g, int img_i0, int img_n, int halo)
{
    const int lo = max(0, g.i_offset - halo), hi = min(g.Nx_global, g.i_offset + g.Nx + halo);
    TAUB_REQUIRE(img_i0 <= lo && img_i0 + img_n >= hi,
                 "image planes [%d, %d) do not cover the required [%d, %d)", img_i0, img_i0 + img_n, lo, hi);
    TAUB_REQUIRE(img_i0 >= 0 && img_i0 + img_n <= g.Nx_global, "image planes outside the volume");
    return TAUB_OK;
}

static ImgView make_view(const taub_geom &g, const uint8_t *img, int img_i0, int img_n)
{
    ImgView v;
    v.img = img;
    v.i0 = img_i0;
    v.n = img_n;
    v.Ny = g.Ny;
    v.Nz = g.Nz;
    v.Nx_global = g.Nx_global;
    v.periodic = g.periodic;
    v.image_stride = (int64_t)img_n * g.Ny * g.Nz;
    return v;
}

static int grid_for(int64_t items, int block)
{
    int64_t blocks = ceil_div64(items, block);
    const int64_t cap = 148 * 64;  // grid-stride loop beyond this
    return (int)(blocks < cap ? blocks : cap);
}

}  // namespace taub

using namespace taub;

extern "C" {

int taub_init_binary(const taub_problem *p, const uint8_t *img, int img_i0, int img_n,
                     const float *vec, void *stream)
{
    TAUB_REQUIRE(p && img && vec, "taub_init_binary: null pointer");
    TAUB_REQUIRE(p->kind == TAUB_BINARY && p->field[0] && p->field[1] && p->codes,
                 "taub_init_binary: problem is not a bound binary problem");
    const taub_geom &g = p->g;
    if (int rc = check_img_cover(g, img_i0, img_n, G + 1)) return rc;
    const int64_t total = (int64_t)taub_codes_elems(&g);
    init_binary_kernel<<<grid_for(total, 256), 256, 0, (cudaStream_t)stream>>>(
        g, make_view(g, img, img_i0, img_n), vec, p->field[0], p->field[1], p->codes);
    TAUB_CUDA(cudaGetLastError());
    count_launch();
    return TAUB_OK;
}

int taub_init_multiphase(const taub_problem *p, const uint8_t *img, int img_i0, int img_n,
                         const uint8_t *map256, const float *cond, const float *vec, void *stream)
{
    TAUB_REQUIRE(p && img && vec && map256 && cond, "taub_init_multiphase: null pointer");
    TAUB_REQUIRE(p->kind == TAUB_MULTIPHASE && p->field[0] && p->field[1] && p->labels && p->lut,
                 "taub_init_multiphase: problem is not a bound multi-phase problem");
    TAUB_REQUIRE(p->L >= 1 && p->L <= TAUB_MAX_LABELS, "taub_init_multiphase: L=%d outside [1, %d]",
                 p->L, TAUB_MAX_LABELS);
    const taub_geom &g = p->g;
    if (int rc = check_img_cover(g, img_i0, img_n, G)) return rc;
    const int64_t total = (int64_t)taub_codes_elems(&g);
    init_multi_kernel<<<grid_for(total, 256), 256, 0, (cudaStream_t)stream>>>(
        g, make_view(g, img, img_i0, img_n), p->L, map256, cond, vec, p->field[0], p->field[1], p->labels);
    TAUB_CUDA(cudaGetLastError());
    count_launch();
    return TAUB_OK;
}

int taub_multiphase_keys(const taub_problem *p, int i_lo, int i_hi, int32_t *keys, void *stream)
{
    TAUB_REQUIRE(p && keys && p->labels, "taub_multiphase_keys: null pointer");
    TAUB_REQUIRE(p->L >= 1 && p->L <= 15, "taub_multiphase_keys: needs at most 15 phases (got %d)", p->L);
    const taub_geom &g = p->g;
    TAUB_REQUIRE(i_lo >= -(G - 1) && i_hi <= g.Nx + (G - 1) && i_lo < i_hi && i_lo + g.i_offset >= 0 &&
                     i_hi + g.i_offset <= g.Nx_global,
                 "taub_multiphase_keys: planes [%d, %d) are not voxel planes held by this slab", i_lo, i_hi);
    const int64_t total = (int64_t)g.bs * (i_hi - i_lo) * g.Ny * g.Nz;
    multiphase_keys_kernel<<<grid_for(total, 256), 256, 0, (cudaStream_t)stream>>>(g, p->labels, keys, i_lo, i_hi - i_lo);
    TAUB_CUDA(cudaGetLastError());
    count_launch();
    return TAUB_OK;
}

size_t taub_class_ws_bytes(void) { return sizeof(ClassWs); }

int taub_class_count(const taub_problem *p, int i_lo, int i_hi, void *ws, void *stream)
{
    TAUB_REQUIRE(p && ws && p->labels, "taub_class_count: null pointer");
    TAUB_REQUIRE(p->L >= 1 && p->L <= 15, "taub_class_count: needs at most 15 phases (got %d)", p->L);
    const taub_geom &g = p->g;
    TAUB_REQUIRE(i_lo >= -(G - 1) && i_hi <= g.Nx + (G - 1) && i_lo < i_hi && i_lo + g.i_offset >= 0 &&
                     i_hi + g.i_offset <= g.Nx_global,
                 "taub_class_count: planes [%d, %d) are not voxel planes held by this slab", i_lo, i_hi);
    cudaStream_t s = (cudaStream_t)stream;
    ClassWs *w = (ClassWs *)ws;
    TAUB_CUDA(cudaMemsetAsync(w, 0, offsetof(ClassWs, keys), s));
    TAUB_CUDA(cudaMemsetAsync(w->keys, 0xff, sizeof(w->keys), s));
    TAUB_CUDA(cudaMemsetAsync(w->counts, 0, sizeof(w->counts), s));
    const int cap = CLS_CAP;
    TAUB_CUDA(cudaMemcpyAsync(&w->capacity, &cap, sizeof(int), cudaMemcpyHostToDevice, s));
    const int64_t total = (int64_t)g.bs * (i_hi - i_lo) * g.Ny * g.Nz;
    class_count_kernel<<<grid_for(total, 256), 256, 0, s>>>(g, p->labels, w, i_lo, i_hi - i_lo);
    TAUB_CUDA(cudaGetLastError());
    count_launch();
    return TAUB_OK;
}

int taub_class_assign(const taub_problem *p, int i_lo, int i_hi, const void *ws, const uint16_t *slot_class, int inert,
                      uint16_t *classes, void *stream)
{
    TAUB_REQUIRE(p && ws && slot_class && classes && p->labels, "taub_class_assign: null pointer");
    TAUB_REQUIRE(inert >= 0 && inert <= 65535, "taub_class_assign: inert class id %d does not fit uint16", inert);
    const taub_geom &g = p->g;
    const int64_t total = (int64_t)taub_codes_elems(&g);
    class_assign_kernel<<<grid_for(total, 256), 256, 0, (cudaStream_t)stream>>>(g, p->labels, (const ClassWs *)ws, slot_class,
                                                                               (unsigned)inert, classes, i_lo, i_hi);
    TAUB_CUDA(cudaGetLastError());
    count_launch();
    return TAUB_OK;
}

int taub_init_anisotropic(const taub_problem *p, const uint8_t *img, int img_i0, int img_n, const float *vec, void *stream)
{
    TAUB_REQUIRE(p && img && vec, "taub_init_anisotropic: null pointer");
    TAUB_REQUIRE(p->kind == TAUB_ANISOTROPIC && p->field[0] && p->field[1] && p->codes,
                 "taub_init_anisotropic: problem is not a bound anisotropic problem");
    const taub_geom &g = p->g;
    TAUB_REQUIRE(!g.periodic, "taub_init_anisotropic: the anisotropic solver has no periodic variant");
    if (int rc = check_img_cover(g, img_i0, img_n, G + 1)) return rc;
    const int64_t total = (int64_t)taub_codes_elems(&g);
    init_classes_kernel<0><<<grid_for(total, 256), 256, 0, (cudaStream_t)stream>>>(
        g, make_view(g, img, img_i0, img_n), vec, p->field[0], p->field[1], p->codes, 1, -1, nullptr);
    TAUB_CUDA(cudaGetLastError());
    count_launch();
    return TAUB_OK;
}

int taub_init_electrode(const taub_problem *p, const uint8_t *img, int cond_label, int reac_label, const float *vec,
                        int64_t *reac_sums, void *stream)
{
    TAUB_REQUIRE(p && img && vec && reac_sums, "taub_init_electrode: null pointer");
    TAUB_REQUIRE(p->kind == TAUB_MULTIPHASE_CLASS && p->field[0] && p->field[1] && p->codes,
                 "taub_init_electrode: problem is not a bound class problem");
    const taub_geom &g = p->g;
    TAUB_REQUIRE(g.i_offset == 0 && g.Nx == g.Nx_global, "taub_init_electrode: whole volumes only");
    TAUB_REQUIRE((int64_t)g.bs * ELECTRODE_IDS < 65535, "taub_init_electrode: batch of %d images needs too many classes", g.bs);
    cudaStream_t s = (cudaStream_t)stream;
    TAUB_CUDA(cudaMemsetAsync(reac_sums, 0, sizeof(int64_t) * g.bs * g.Nx, s));
    const int64_t total = (int64_t)taub_codes_elems(&g);
    init_classes_kernel<1><<<grid_for(total, 256), 256, 0, s>>>(g, make_view(g, img, 0, g.Nx), vec, p->field[0], p->field[1],
                                                               p->codes, cond_label, reac_label, (unsigned long long *)reac_sums);
    TAUB_CUDA(cudaGetLastError());
    count_launch();
    return TAUB_OK;
}

int taub_plane_counts(const taub_geom *g, const uint8_t *img, int img_i0, int img_n,
                      const uint8_t *sel256, int64_t *counts, int64_t *hist, void *stream)
{
    TAUB_REQUIRE(g && img && sel256 && counts, "taub_plane_counts: null pointer");
    if (int rc = check_img_cover(*g, img_i0, img_n, 0)) return rc;
    cudaStream_t s = (cudaStream_t)stream;
    if (hist) TAUB_CUDA(cudaMemsetAsync(hist, 0, sizeof(int64_t) * 256 * g->bs, s));
    dim3 grid(g->Nx, g->bs);
    plane_counts_kernel<<<grid, 256, 0, s>>>(*g, make_view(*g, img, img_i0, img_n), sel256,
                                             (unsigned long long *)counts, (unsigned long long *)hist);
    TAUB_CUDA(cudaGetLastError());
    count_launch();
    return TAUB_OK;
}

}  // extern "C"
