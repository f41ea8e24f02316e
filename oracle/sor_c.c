/* TEST INFRASTRUCTURE ONLY (see oracle/__init__.py) -- never linked into the product.
 *
 * Plain-C restatement of the reference's hot loop in the reference's own padded layout
 *   field[bs][Nx+2][Ny+2][Nz+2]  (fp32, x = flux direction = slowest dim),
 *   factor[bs][Nx][Ny][Nz], D_x[bs][Nx+1][Ny][Nz], D_y[bs][Nx][Ny+1][Nz], D_z[bs][Nx][Ny][Nz+1].
 * Follows /root/reference/taufactor/taufactor.py:
 *   ghost refresh  :501-505 / :652-656     neighbour sum :95-103 / :606-613
 *   update         :176-181 (sum / factor - f, times omega on the active colour, added to f)
 *   flux           :412-419 / :615-620     plane means   :296, :307
 * Build with  gcc -O2 -ffp-contract=off  (no FMA contraction, IEEE division) -- see Makefile.
 * Each active voxel only reads opposite-colour neighbours (ghost cells are never written inside
 * a sweep), so the in-place loop equals the reference's whole-array expression bit for bit.
 */
#include <stddef.h>
#include <stdint.h>

#define IDX(b, i, j, k) ((((size_t)(b) * PX + (i)) * PY + (j)) * PZ + (k))

void orc_refresh_ghosts(float *f, int bs, int Nx, int Ny, int Nz)
{
    const size_t PX = Nx + 2, PY = Ny + 2, PZ = Nz + 2;
    for (int b = 0; b < bs; ++b)
        for (int i = 0; i < (int)PX; ++i) {
            for (int k = 0; k < (int)PZ; ++k) {            /* y ghosts first */
                f[IDX(b, i, 0, k)] = f[IDX(b, i, PY - 2, k)];
                f[IDX(b, i, PY - 1, k)] = f[IDX(b, i, 1, k)];
            }
            for (int j = 0; j < (int)PY; ++j) {            /* then z ghosts (all rows) */
                f[IDX(b, i, j, 0)] = f[IDX(b, i, j, PZ - 2)];
                f[IDX(b, i, j, PZ - 1)] = f[IDX(b, i, j, 1)];
            }
        }
}

/* n reference iterations starting at iteration counter iter0 (colour = iter % 2 on the
 * 0-based interior indices). Dx == NULL selects the binary solvers. */
void orc_half_sweep_range(float *f, const float *factor, const float *Dx, const float *Dy,
                          const float *Dz, int bs, int Nx, int Ny, int Nz, float omega, int colour,
                          int a0, int a1);

void orc_sweeps(float *f, const float *factor, const float *Dx, const float *Dy, const float *Dz,
                int bs, int Nx, int Ny, int Nz, int periodic, float omega, long iter0, int n)
{
    for (int it = 0; it < n; ++it) {
        if (periodic)
            orc_refresh_ghosts(f, bs, Nx, Ny, Nz);
        orc_half_sweep_range(f, factor, Dx, Dy, Dz, bs, Nx, Ny, Nz, omega, (int)((iter0 + it) & 1), 0, Nx);
    }
}

/* One colour over the x planes [a0, a1) of every image: the unit Python threads split. */
void orc_half_sweep_range(float *f, const float *factor, const float *Dx, const float *Dy,
                          const float *Dz, int bs, int Nx, int Ny, int Nz, float omega, int colour,
                          int a0, int a1)
{
    const size_t PX = Nx + 2, PY = Ny + 2, PZ = Nz + 2;
    {
        for (int b = 0; b < bs; ++b)
            for (int a = a0; a < a1; ++a)
                for (int c = 0; c < Ny; ++c) {
                    const size_t fo = ((size_t)b * Nx + a) * Ny + c;   /* factor row */
                    for (int d = (a + c + colour) & 1; d < Nz; d += 2) {
                        const size_t p = IDX(b, a + 1, c + 1, d + 1);
                        float s;
                        if (!Dx) {
                            s = f[p + PY * PZ] + f[p - PY * PZ];
                            s = s + f[p + PZ];
                            s = s + f[p - PZ];
                            s = s + f[p + 1];
                            s = s + f[p - 1];
                        } else {
                            const size_t ox = (((size_t)b * (Nx + 1) + a) * Ny + c) * Nz + d;
                            const size_t oy = (((size_t)b * Nx + a) * (Ny + 1) + c) * Nz + d;
                            const size_t oz = (((size_t)b * Nx + a) * Ny + c) * (Nz + 1) + d;
                            s = f[p + PY * PZ] * Dx[ox + (size_t)Ny * Nz] + f[p - PY * PZ] * Dx[ox];
                            s = s + f[p + PZ] * Dy[oy + Nz];
                            s = s + f[p - PZ] * Dy[oy];
                            s = s + f[p + 1] * Dz[oz + 1];
                            s = s + f[p - 1] * Dz[oz];
                        }
                        float inc = s / factor[fo * Nz + d];
                        inc = inc - f[p];
                        inc = inc * omega;
                        f[p] = f[p] + inc;
                    }
                }
    }
}

/* AnisotropicSolver (taufactor.py:473-478): s = ((x+ + x-) + Ky*(y+ + y-)) + Kz*(z+ + z-), Ky / Kz already
 * rounded to fp32 (torch multiplies an fp32 tensor by the Python scalar in fp32); update as above, no
 * periodic variant. */
void orc_sweeps_aniso(float *f, const float *factor, int bs, int Nx, int Ny, int Nz, float omega, float Ky,
                      float Kz, long iter0, int n)
{
    const size_t PX = Nx + 2, PY = Ny + 2, PZ = Nz + 2;
    for (int it = 0; it < n; ++it) {
        const int colour = (int)((iter0 + it) & 1);
        for (int b = 0; b < bs; ++b)
            for (int a = 0; a < Nx; ++a)
                for (int c = 0; c < Ny; ++c) {
                    const size_t fo = ((size_t)b * Nx + a) * Ny + c;
                    for (int d = (a + c + colour) & 1; d < Nz; d += 2) {
                        const size_t p = IDX(b, a + 1, c + 1, d + 1);
                        float s = f[p + PY * PZ] + f[p - PY * PZ];
                        float t = f[p + PZ] + f[p - PZ];
                        t = Ky * t;
                        s = s + t;
                        t = f[p + 1] + f[p - 1];
                        t = Kz * t;
                        s = s + t;
                        float inc = s / factor[fo * Nz + d];
                        inc = inc - f[p];
                        inc = inc * omega;
                        f[p] = f[p] + inc;
                    }
                }
    }
}

/* Per-x-plane sums (fp64): flux_sum[b][i] over faces i|i+1, i = 0..Nx-2, and field_sum[b][i]. */
void orc_plane_sums(const float *f, const float *factor, const float *Dx, int bs, int Nx, int Ny,
                    int Nz, double *flux_sum, double *field_sum)
{
    const size_t PX = Nx + 2, PY = Ny + 2, PZ = Nz + 2;
    for (int b = 0; b < bs; ++b)
        for (int a = 0; a < Nx; ++a) {
            double fs = 0.0, cs = 0.0;
            for (int c = 0; c < Ny; ++c)
                for (int d = 0; d < Nz; ++d) {
                    const size_t p = IDX(b, a + 1, c + 1, d + 1);
                    cs += (double)f[p];
                    if (a + 1 < Nx) {
                        float v = f[p + PY * PZ] - f[p];
                        if (!Dx) {
                            const size_t fo = (((size_t)b * Nx + a) * Ny + c) * Nz + d;
                            if (factor[fo] > 8.0f || factor[fo + (size_t)Ny * Nz] > 8.0f)
                                v = 0.0f;
                        } else {
                            v = Dx[(((size_t)b * (Nx + 1) + a + 1) * Ny + c) * Nz + d] * v;
                        }
                        fs += (double)v;
                    }
                }
            field_sum[(size_t)b * Nx + a] = cs;
            if (a + 1 < Nx)
                flux_sum[(size_t)b * (Nx - 1) + a] = fs;
        }
}
