// Test harness (never part of the product library): runs the host instantiation of taub::refresh_plane_v2 --
// the function the device kernel refresh_ghosts_v2_kernel calls -- over a whole storage array, emulating
// `nthreads` callers per plane one after the other.
#include <cuda_runtime.h>
#include <stdint.h>

#include "taub_refresh.cuh"

extern "C" void refresh_field_host(const taub_geom *g, float *field, int nthreads)
{
    for (int b = 0; b < g->bs; ++b)
        for (int p = 0; p < g->planes; ++p) {
            float *plane = field + (int64_t)b * g->image_stride + (int64_t)p * g->plane_stride;
            for (int tid = nthreads - 1; tid >= 0; --tid)      // any order must do: run them backwards
                taub::refresh_plane_v2(*g, plane, tid, nthreads);
        }
}
