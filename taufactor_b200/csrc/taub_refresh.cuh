// taub_refresh.cuh -- periodic ghost frame of ONE storage plane (taufactor.py:501-505 / :652-656), written so
// that the same function runs on the device (refresh_ghosts_v2_kernel) and on the host (tests/csrc/
// refresh_host.cu: the index logic is checked on the CPU against a NumPy wrap-pad).
//
// Ghost rows [0,G) u [G+Ny, rows) := image of the wrapped interior row over the frame columns [COL0-G, COL0+Nz+G)
// (0 in the padding columns outside the frame); ghost column pairs {COL0-2, COL0-1} and {COL0+Nz, COL0+Nz+1} of the
// interior rows := image of the wrapped interior column.  Reads interior cells only and writes ghost cells only,
// so the work items are independent: `nthreads` callers tid = 0..nthreads-1 may run in any order.
// Differences to refresh_ghosts_kernel: no integer division per item (nested loops; the wrap of an index in
// [-G, n+G) is one compare-and-add), float4 copies for the groups that lie inside the interior columns, one CTA
// (pair) per plane instead of seven.
#ifndef TAUB_REFRESH_CUH
#define TAUB_REFRESH_CUH

#include "taub200.h"

namespace taub {

// a in [-G, n+G) -> [0, n); general modulo for extents smaller than the ghost width
static inline __host__ __device__ int wrap_near(int a, int n)
{
    if (n >= TAUB_GHOST) return a < 0 ? a + n : (a >= n ? a - n : a);
    a %= n;
    return a < 0 ? a + n : a;
}

static inline __host__ __device__ void refresh_plane_v2(const taub_geom &g, float *plane, int tid, int nthreads)
{
    constexpr int G_ = TAUB_GHOST, C0 = TAUB_COL0;
    const int PG = g.pitch >> 2;
#pragma unroll
    for (int r = 0; r < 2 * G_; ++r) {
        const int jr = r < G_ ? r : g.Ny + r;                 // r in [G, 2G) -> rows [G+Ny, 2G+Ny)
        const float *src = plane + (int64_t)(G_ + wrap_near(jr - G_, g.Ny)) * g.pitch;
        float *dst = plane + (int64_t)jr * g.pitch;
        for (int grp = tid; grp < PG; grp += nthreads) {
            const int c = 4 * grp;
            float4 v;
            if (c >= C0 && c + 3 < C0 + g.Nz) {               // wholly inside the interior columns: plain copy
                v = *reinterpret_cast<const float4 *>(src + c);
            } else {
                float w[4];
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    const bool frame = (c + q >= C0 - G_) && (c + q < C0 + g.Nz + G_);
                    w[q] = frame ? src[C0 + wrap_near(c + q - C0, g.Nz)] : 0.0f;
                }
                v.x = w[0]; v.y = w[1]; v.z = w[2]; v.w = w[3];
            }
            *reinterpret_cast<float4 *>(dst + c) = v;
        }
    }
    for (int u = tid; u < 2 * g.Ny; u += nthreads) {
        const int side = u & 1;
        float *row = plane + (int64_t)(G_ + (u >> 1)) * g.pitch;
        const int c = side ? C0 + g.Nz : C0 - G_;             // first of the two ghost columns
        const float v0 = row[C0 + wrap_near(c - C0, g.Nz)], v1 = row[C0 + wrap_near(c + 1 - C0, g.Nz)];
        row[c] = v0;
        row[c + 1] = v1;
    }
}

}  // namespace taub
#endif
