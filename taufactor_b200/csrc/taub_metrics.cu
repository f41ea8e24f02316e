// taub_metrics.cu -- the periodic flux / convergence reduction.
//
// Replaces vertical_flux (taufactor.py:412-419 binary: f[i+1]-f[i] zeroed where either voxel has
// factor = inf; :615-620 multi-phase: D_x * (f[i+1]-f[i])) plus the two full-volume torch.mean
// reductions of compute_metrics (:296 flux_1d, :307 mean field per slice): one fused pass that
// reads the field once, warp-shuffle + block reduction in fp64, deterministic two-stage sum.
#include "taub_common.cuh"

namespace taub {

constexpr int SUM_THREADS = 256;

static int sum_chunks(const taub_geom &g)
{
    // rows per block: enough blocks to fill the GPU, few enough partials to keep stage 2 trivial
    int chunks = ceil_div(g.Ny, 16);
    return chunks > 64 ? 64 : (chunks < 1 ? 1 : chunks);
}

template <int KIND>
__global__ void __launch_bounds__(SUM_THREADS)
plane_sums_kernel(taub_geom g, const float *__restrict__ f, const uint16_t *__restrict__ codes,
                  const uint8_t *__restrict__ labels, const float *__restrict__ lut, int L,
                  int n_flux, double2 *__restrict__ partial, const int *__restrict__ stop)
{
    if (stop && *stop) return;
    constexpr bool MULTI = (KIND == TAUB_MULTIPHASE);
    extern __shared__ float s_lut[];
    __shared__ double s_red[2][SUM_THREADS / 32];
    if (MULTI) {
        for (int t = threadIdx.x; t < (L + 1) * (L + 1); t += blockDim.x) s_lut[t] = lut[t];
        __syncthreads();
    }
    const int chunk = blockIdx.x, nchunks = gridDim.x;
    const int il = blockIdx.y, b = blockIdx.z;
    const int j0 = (int)((int64_t)g.Ny * chunk / nchunks), j1 = (int)((int64_t)g.Ny * (chunk + 1) / nchunks);
    const int ng = interior_groups(g.Nz);
    const bool has_flux = il < n_flux;
    const int64_t base = (int64_t)b * g.image_stride + (int64_t)(il + G) * g.plane_stride;
    double flux = 0.0, fsum = 0.0;
    const int items = (j1 - j0) * ng;
    for (int t = threadIdx.x; t < items; t += blockDim.x) {
        const int r = t / ng, grp = t - r * ng;
        const int64_t o = base + (int64_t)(j0 + r + G) * g.pitch + COL0 + 4 * grp;
        const float4 a4 = *reinterpret_cast<const float4 *>(f + o);
        const float a[4] = {a4.x, a4.y, a4.z, a4.w};
        const int valid = min(4, g.Nz - 4 * grp);
        float n[4] = {0.f, 0.f, 0.f, 0.f};
        unsigned ca = 0, cn = 0;
        uint32_t la = 0, ln = 0;
        if (has_flux) {
            const float4 n4 = *reinterpret_cast<const float4 *>(f + o + g.plane_stride);
            n[0] = n4.x; n[1] = n4.y; n[2] = n4.z; n[3] = n4.w;
            if (KIND == TAUB_BINARY) {
                ca = codes[o >> 2];
                cn = codes[(o + g.plane_stride) >> 2];
            } else if (MULTI) {
                la = *reinterpret_cast<const uint32_t *>(labels + o);
                ln = *reinterpret_cast<const uint32_t *>(labels + o + g.plane_stride);
            }
        }
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            if (q < valid) {
                fsum += (double)a[q];
                if (has_flux) {
                    float v = __fsub_rn(n[q], a[q]);
                    if (KIND == TAUB_BINARY) {
                        // factor > 8 (inf) on either side zeroes the flux (ref:417-418): codes 1..8 are finite
                        const bool open = (((ca >> (4 * q)) & 15u) - 1u) < 8u && (((cn >> (4 * q)) & 15u) - 1u) < 8u;
                        v = open ? v : 0.0f;
                    } else if (KIND == TAUB_MULTIPHASE_CLASS) {
                        // D_x face conductance towards plane i+1 = first entry of the voxel's class row
                        v = __fmul_rn(__ldg(lut + 8 * (int)codes[o + q]), v);
                    } else if (KIND == TAUB_ANISOTROPIC) {
                        // the reference's factor > 8 test (:417-418) on the weighted prefactors of both voxels
                        // (b = 0 encodes inf); the prefactor may legitimately exceed 8 here
                        const float fa = __ldg(lut + 2 * (int)codes[o + q]);
                        const float fn = __ldg(lut + 2 * (int)codes[o + g.plane_stride + q]);
                        v = (fa == 0.0f || fa > 8.0f || fn == 0.0f || fn > 8.0f) ? 0.0f : v;
                    } else {
                        v = __fmul_rn(s_lut[((la >> (8 * q)) & 255u) * (L + 1) + ((ln >> (8 * q)) & 255u)], v);
                    }
                    flux += (double)v;
                }
            }
        }
    }
    for (int o = 16; o > 0; o >>= 1) {
        flux += __shfl_xor_sync(0xffffffffu, flux, o);
        fsum += __shfl_xor_sync(0xffffffffu, fsum, o);
    }
    const int w = threadIdx.x >> 5;
    if ((threadIdx.x & 31) == 0) {
        s_red[0][w] = flux;
        s_red[1][w] = fsum;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        double a = 0.0, c = 0.0;
        for (int k = 0; k < SUM_THREADS / 32; ++k) {
            a += s_red[0][k];
            c += s_red[1][k];
        }
        partial[((int64_t)b * g.Nx + il) * nchunks + chunk] = make_double2(a, c);
    }
}

__global__ void finalize_means_kernel(taub_geom g, const double2 *__restrict__ partial, int nchunks,
                                      int n_flux, float *__restrict__ flux_mean, float *__restrict__ field_mean,
                                      const int *__restrict__ stop)
{
    if (stop && *stop) return;
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= g.bs * g.Nx) return;
    const int b = t / g.Nx, il = t - b * g.Nx;
    double a = 0.0, c = 0.0;
    for (int k = 0; k < nchunks; ++k) {
        const double2 p = partial[(int64_t)t * nchunks + k];
        a += p.x;
        c += p.y;
    }
    const double n = (double)g.Ny * (double)g.Nz;
    field_mean[t] = (float)(c / n);
    if (il < n_flux) flux_mean[(int64_t)b * n_flux + il] = (float)(a / n);
}

// NumPy's float32 add.reduce along a contiguous axis (loops_utils.h.src, pairwise sum): plain loop
// below 8 elements, 8 interleaved accumulators up to 128, recursive halving (multiples of 8) above.
__device__ float np_pairwise_sum(const float *a, int n)
{
    if (n < 8) {
        float res = 0.0f;
        for (int i = 0; i < n; ++i) res = __fadd_rn(res, a[i]);
        return res;
    }
    if (n <= 128) {
        float r[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) r[j] = a[j];
        int i = 8;
        for (; i < n - (n % 8); i += 8) {
#pragma unroll
            for (int j = 0; j < 8; ++j) r[j] = __fadd_rn(r[j], a[i + j]);
        }
        float res = __fadd_rn(__fadd_rn(__fadd_rn(r[0], r[1]), __fadd_rn(r[2], r[3])),
                              __fadd_rn(__fadd_rn(r[4], r[5]), __fadd_rn(r[6], r[7])));
        for (; i < n; ++i) res = __fadd_rn(res, a[i]);
        return res;
    }
    int n2 = n / 2;
    n2 -= n2 % 8;
    return __fadd_rn(np_pairwise_sum(a, n2), np_pairwise_sum(a + n2, n - n2));
}

// The reference's compute_metrics scalars (taufactor.py:297-305) and check_convergence decision
// (:143-153) in the same float32 arithmetic NumPy uses on the host.  One thread per image, one block.
__global__ void stop_rule_kernel(int bs, int Nx, const float *__restrict__ flux_mean, const double *__restrict__ D_mean,
                                 float *__restrict__ old_tau, float conv_crit, float *__restrict__ record,
                                 int *__restrict__ stop)
{
    __shared__ int s_all_conv, s_needs_host;
    __shared__ float s_tau_err;
    if (*stop) {                        // a previous check already stopped the solve
        if (threadIdx.x == 0) record[0] = 3.0f;
        return;
    }
    if (threadIdx.x == 0) {
        s_all_conv = 1;
        s_needs_host = 0;
        s_tau_err = 0.0f;
    }
    __syncthreads();
    const int n = Nx - 1;
    for (int b = threadIdx.x; b < bs; b += blockDim.x) {
        const float *fl = flux_mean + (size_t)b * n;
        float fmax = fl[0], fmin = fl[0];
        for (int i = 1; i < n; ++i) {
            fmax = fmaxf(fmax, fl[i]);   // NaN handling differs from np.max only when the field is NaN
            fmin = fminf(fmin, fl[i]);
        }
        const float mean_fl = __fdiv_rn(np_pairwise_sum(fl, n), (float)n);
        float rel = (fmax != 0.0f) ? __fdiv_rn(__fsub_rn(fmax, fmin), fmax) : __int_as_float(0x7fc00000);
        const float D_rel = __fmul_rn(mean_fl, (float)Nx);                     // / abs(top_bc - bot_bc) == 1
        // np.divide(D_mean, D_rel, out=float32): D_mean is float32 (binary) or float64 (multi-phase); the
        // double quotient rounded to float equals the float32 quotient of float32 operands (53 >= 2*24+2)
        const float tau = (D_rel != 0.0f) ? (float)(D_mean[b] / (double)D_rel) : __int_as_float(0x7fc00000);
        if (fmin == 0.0f || fmax == 0.0f || mean_fl == 0.0f) s_needs_host = 1;  // percolation check, host only
        if (mean_fl != mean_fl) rel = 0.0f;                                     // NaN counts as converged
        record[2 + b] = tau;
        record[2 + bs + b] = rel;
        if (!(rel < conv_crit)) s_all_conv = 0;
        const float err = fabsf(__fsub_rn(tau, old_tau[b]));
        if (!(err < 2e-3f)) atomicExch(reinterpret_cast<int *>(&s_tau_err), __float_as_int(1.0f));  // any failure
    }
    __syncthreads();
    const bool converged = s_all_conv && !(s_tau_err > 0.0f);
    if (!s_needs_host && !converged)
        for (int b = threadIdx.x; b < bs; b += blockDim.x) old_tau[b] = record[2 + b];   // ref:144,149
    if (threadIdx.x == 0) {
        const int status = s_needs_host ? 2 : (converged ? 1 : 0);
        record[0] = (float)status;
        record[1] = 0.0f;
        if (status) *stop = status;
    }
}

}  // namespace taub

using namespace taub;

extern "C" {

size_t taub_sums_ws_bytes(const taub_geom *g)
{
    return sizeof(double2) * (size_t)g->bs * g->Nx * sum_chunks(*g);
}

static int plane_means(const taub_problem *p, void *workspace, float *flux_mean, float *field_mean,
                       const int *stop, void *stream);

int taub_plane_means(const taub_problem *p, void *workspace, float *flux_mean, float *field_mean,
                     void *stream)
{
    return plane_means(p, workspace, flux_mean, field_mean, p ? p->stop : nullptr, stream);
}

int taub_stop_rule_async(int bs, int Nx_global, const float *flux_mean, const double *D_mean, float *old_tau,
                         float conv_crit, float *record, int32_t *stop, void *stream)
{
    TAUB_REQUIRE(flux_mean && D_mean && old_tau && record && stop, "taub_stop_rule_async: null pointer");
    TAUB_REQUIRE(bs >= 1 && Nx_global >= 2, "taub_stop_rule_async: needs bs >= 1 and Nx_global >= 2");
    stop_rule_kernel<<<1, 64, 0, (cudaStream_t)stream>>>(bs, Nx_global, flux_mean, D_mean, old_tau, conv_crit, record, stop);
    TAUB_CUDA(cudaGetLastError());
    count_launch();
    return TAUB_OK;
}

int taub_check_async(const taub_problem *p, void *workspace, float *flux_mean, float *field_mean,
                     const double *D_mean, float *old_tau, float conv_crit, float *record, void *stream)
{
    TAUB_REQUIRE(p && D_mean && old_tau && record, "taub_check_async: null pointer");
    TAUB_REQUIRE(p->stop != nullptr, "taub_check_async: the problem has no stop flag");
    TAUB_REQUIRE(p->g.i_offset == 0 && p->g.Nx == p->g.Nx_global && p->g.Nx >= 2,
                 "taub_check_async needs the whole volume on this device and Nx >= 2");
    if (int rc = plane_means(p, workspace, flux_mean, field_mean, p->stop, stream)) return rc;
    stop_rule_kernel<<<1, 64, 0, (cudaStream_t)stream>>>(p->g.bs, p->g.Nx, flux_mean, D_mean, old_tau, conv_crit,
                                                        record, p->stop);
    TAUB_CUDA(cudaGetLastError());
    count_launch();
    return TAUB_OK;
}

static int plane_means(const taub_problem *p, void *workspace, float *flux_mean, float *field_mean,
                       const int *stop, void *stream)
{
    TAUB_REQUIRE(p && workspace && flux_mean && field_mean, "taub_plane_means: null pointer");
    const taub_geom &g = p->g;
    TAUB_REQUIRE(g.bs <= 65535 && g.Nx <= 65535, "taub_plane_means: bs / Nx above 65535");
    const int has_next = (g.i_offset + g.Nx < g.Nx_global) ? 1 : 0;
    const int n_flux = g.Nx - 1 + has_next;
    const int nchunks = sum_chunks(g);
    const float *f = p->field[p->cur];
    cudaStream_t s = (cudaStream_t)stream;
    dim3 grid(nchunks, g.Nx, g.bs);
    if (p->kind == TAUB_BINARY) {
        plane_sums_kernel<TAUB_BINARY><<<grid, SUM_THREADS, 0, s>>>(g, f, p->codes, nullptr, nullptr, 0, n_flux,
                                                                    (double2 *)workspace, stop);
    } else if (p->kind == TAUB_MULTIPHASE_CLASS) {
        plane_sums_kernel<TAUB_MULTIPHASE_CLASS><<<grid, SUM_THREADS, 0, s>>>(g, f, p->codes, nullptr, p->lut, p->L, n_flux,
                                                                              (double2 *)workspace, stop);
    } else if (p->kind == TAUB_ANISOTROPIC) {
        plane_sums_kernel<TAUB_ANISOTROPIC><<<grid, SUM_THREADS, 0, s>>>(g, f, p->codes, nullptr, p->lut, 0, n_flux,
                                                                         (double2 *)workspace, stop);
    } else {
        const size_t smem = sizeof(float) * (p->L + 1) * (p->L + 1);
        plane_sums_kernel<TAUB_MULTIPHASE><<<grid, SUM_THREADS, smem, s>>>(g, f, nullptr, p->labels, p->lut, p->L,
                                                                n_flux, (double2 *)workspace, stop);
    }
    TAUB_CUDA(cudaGetLastError());
    count_launch();
    finalize_means_kernel<<<ceil_div(g.bs * g.Nx, 128), 128, 0, s>>>(g, (const double2 *)workspace, nchunks,
                                                                    n_flux, flux_mean, field_mean, stop);
    TAUB_CUDA(cudaGetLastError());
    count_launch();
    return TAUB_OK;
}

}  // extern "C"
