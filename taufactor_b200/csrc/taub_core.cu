// taub_core.cu -- error channel, device facts, slab-storage geometry.
#include <stdarg.h>
#include <string.h>

#include <atomic>

#include "taub_common.cuh"

namespace taub {
static thread_local char g_err[512] = "";

void set_error(const char *fmt, ...)
{
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}
static std::atomic<unsigned long long> g_launches{0};
void count_launch(int n) { g_launches.fetch_add((unsigned long long)n, std::memory_order_relaxed); }
}  // namespace taub

extern "C" {

int taub_abi_version(void) { return TAUB_ABI_VERSION; }

unsigned long long taub_launch_count(void) { return taub::g_launches.load(std::memory_order_relaxed); }

const char *taub_last_error(void) { return taub::g_err; }

int taub_device_info(int *sm_count, int *cc_major, int *cc_minor, int *runtime_version)
{
    int dev = 0;
    TAUB_CUDA(cudaGetDevice(&dev));
    cudaDeviceProp prop;
    TAUB_CUDA(cudaGetDeviceProperties(&prop, dev));
    if (sm_count) *sm_count = prop.multiProcessorCount;
    if (cc_major) *cc_major = prop.major;
    if (cc_minor) *cc_minor = prop.minor;
    if (runtime_version) TAUB_CUDA(cudaRuntimeGetVersion(runtime_version));
    return TAUB_OK;
}

int taub_set_device(int ordinal)
{
    TAUB_CUDA(cudaSetDevice(ordinal));
    return TAUB_OK;
}

int taub_geom_init(taub_geom *g, int bs, int Nx_local, int Ny, int Nz, int Nx_global,
                   int i_offset, int periodic)
{
    TAUB_REQUIRE(g != nullptr, "taub_geom_init: null geometry");
    TAUB_REQUIRE(bs >= 1 && Nx_local >= 1 && Ny >= 1 && Nz >= 1,
                 "taub_geom_init: extents must be >= 1 (bs=%d Nx=%d Ny=%d Nz=%d)", bs, Nx_local, Ny, Nz);
    TAUB_REQUIRE(Nx_global >= Nx_local && i_offset >= 0 && i_offset + Nx_local <= Nx_global,
                 "taub_geom_init: slab [%d, %d) outside the volume of %d planes", i_offset,
                 i_offset + Nx_local, Nx_global);
    memset(g, 0, sizeof(*g));
    g->bs = bs;
    g->Nx = Nx_local;
    g->Ny = Ny;
    g->Nz = Nz;
    g->Nx_global = Nx_global;
    g->i_offset = i_offset;
    g->periodic = periodic ? 1 : 0;
    g->planes = Nx_local + 2 * TAUB_GHOST;
    g->rows = Ny + 2 * TAUB_GHOST;
    // interior at column 4 (16-byte aligned), z ghosts either side, then room for the float4
    // group that holds the last interior voxel plus its right-hand neighbour; rows are multiples of
    // 128 bytes (full cache lines; the uint16 code rows are then multiples of 16 bytes, which TMA needs).
    g->pitch = ((Nz + 8 + 31) / 32) * 32;
    g->plane_stride = (int64_t)g->rows * g->pitch;
    g->image_stride = (int64_t)g->planes * g->plane_stride;
    return TAUB_OK;
}

size_t taub_field_elems(const taub_geom *g) { return (size_t)g->bs * (size_t)g->image_stride; }

size_t taub_codes_elems(const taub_geom *g) { return taub_field_elems(g) / 4; }

}  // extern "C"
