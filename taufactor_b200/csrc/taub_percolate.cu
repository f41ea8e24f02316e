// Percolation check on the device: does a 6-connected cluster of the conductive mask join the first and
// the last x plane?  Replaces the SciPy labelling of the whole volume that the reference runs on the
// host whenever a slice flux is exactly 0 (taufactor.py:318-327 -> metrics/connectivity.py:138-213,
// extract_through_feature(mask, 1, 'x'): non-periodic, connectivity 1, spanning along x).
//
// Method: flood fill from plane 0 by directional marches.  One round = a forward + backward march along
// x, then y, then z; a march carries the "reached" state along a whole grid line in one kernel, so a
// round advances the front through any number of straight segments.  Rounds repeat until no byte flips.
// Dense byte arrays mask / reach [bs][Nx][Ny][Nz]; this is a rare-event path, HBM-bound byte work.
#include "taub_common.cuh"

namespace taub {

// One thread per grid line.  (n, stride) describe the marched axis; a line starts at base.
__device__ __forceinline__ bool march_line(const uint8_t *__restrict__ mask, uint8_t *__restrict__ reach,
                                           int64_t base, int n, int64_t stride)
{
    bool changed = false;
    uint8_t carry = 0;
    for (int t = 0; t < n; ++t) {            // forward
        const int64_t o = base + t * stride;
        const uint8_t m = mask[o], r = reach[o];
        const uint8_t now = m ? (uint8_t)(r | carry) : (uint8_t)0;
        if (now != r) {
            reach[o] = now;
            changed = true;
        }
        carry = now;
    }
    carry = 0;
    for (int t = n - 1; t >= 0; --t) {       // backward
        const int64_t o = base + t * stride;
        const uint8_t m = mask[o], r = reach[o];
        const uint8_t now = m ? (uint8_t)(r | carry) : (uint8_t)0;
        if (now != r) {
            reach[o] = now;
            changed = true;
        }
        carry = now;
    }
    return changed;
}

// axis 0: lines along x, one per (b, y, z); axis 1: along y, one per (b, x, z); axis 2: along z, one
// per (b, x, y).  Consecutive threads take consecutive z (axes 0, 1) or consecutive y (axis 2).
__global__ void __launch_bounds__(256)
flood_march_kernel(const uint8_t *__restrict__ mask, uint8_t *__restrict__ reach, int bs, int Nx, int Ny, int Nz,
                   int axis, int *__restrict__ changed)
{
    const int64_t sx = (int64_t)Ny * Nz, sy = Nz, img = (int64_t)Nx * sx;
    const int64_t lines = (axis == 0) ? (int64_t)bs * Ny * Nz : (axis == 1) ? (int64_t)bs * Nx * Nz : (int64_t)bs * Nx * Ny;
    bool any = false;
    for (int64_t l = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; l < lines; l += (int64_t)gridDim.x * blockDim.x) {
        if (axis == 0) {
            const int64_t b = l / sx, yz = l - b * sx;
            any |= march_line(mask, reach, b * img + yz, Nx, sx);
        } else if (axis == 1) {
            const int64_t bx = l / Nz, z = l - bx * Nz;
            any |= march_line(mask, reach, bx * sx + z, Ny, sy);
        } else {
            any |= march_line(mask, reach, l * Nz, Nz, 1);
        }
    }
    if (__syncthreads_or(any) && threadIdx.x == 0) *changed = 1;
}

}  // namespace taub

using namespace taub;

extern "C" {

int taub_flood_round(const uint8_t *mask, uint8_t *reach, int bs, int Nx, int Ny, int Nz, int *changed, void *stream)
{
    TAUB_REQUIRE(mask && reach && changed, "taub_flood_round: null pointer");
    TAUB_REQUIRE(bs >= 1 && Nx >= 1 && Ny >= 1 && Nz >= 1, "taub_flood_round: empty volume");
    cudaStream_t s = (cudaStream_t)stream;
    TAUB_CUDA(cudaMemsetAsync(changed, 0, sizeof(int), s));
    for (int axis = 0; axis < 3; ++axis) {
        const int64_t lines = (axis == 0) ? (int64_t)bs * Ny * Nz : (axis == 1) ? (int64_t)bs * Nx * Nz : (int64_t)bs * Nx * Ny;
        const int n = (axis == 0) ? Nx : (axis == 1) ? Ny : Nz;
        if (n < 2) continue;                      // nothing to carry along a line of one voxel
        const int blocks = (int)((lines + 255) / 256);
        flood_march_kernel<<<blocks < 1 ? 1 : blocks, 256, 0, s>>>(mask, reach, bs, Nx, Ny, Nz, axis, changed);
        TAUB_CUDA(cudaGetLastError());
        count_launch();
    }
    return TAUB_OK;
}

}  // extern "C"
