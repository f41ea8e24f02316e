/* taub200.h -- C ABI of libtaub200.so, the B200 (sm_100a) kernels behind the TauFactor
 * steady-state diffusion solve.
 *
 * The reference (tldr-group/taufactor v1.2.1) has NO FFI of its own: its hot path is a chain of
 * eager PyTorch tensor expressions inside taufactor/taufactor.py.  Each entry point below
 * replaces the reference lines cited next to it; the Python classes in taufactor_b200/solvers.py
 * keep the reference's surface (Solver / PeriodicSolver / MultiPhaseSolver /
 * PeriodicMultiPhaseSolver, .solve(), .tau, .D_eff ...) and call these through ctypes.
 *
 * Conventions: extern "C"; plain pointers and sizes only (no torch types); every function
 * returns 0 on success and a negative taub_status otherwise, with a message retrievable from
 * taub_last_error(); nothing throws, nothing calls back into Python.  The library never
 * allocates fields: the caller (PyTorch) owns all device memory and passes raw device pointers
 * plus a cudaStream_t (as void*).  All launches are asynchronous on that stream.
 *
 * Storage layout ("slab storage"): per image
 *     field[planes][rows][pitch]   fp32,  planes = Nx + 2G, rows = Ny + 2G, G = 2
 * voxel (i, j, k) of the local slab lives at plane i+G, row j+G, column k+4 (so interior rows
 * start 16-byte aligned; columns 2,3 and Nz+4,Nz+5 are the z ghosts).  The reference's padded
 * tensor field[bs, Nx+2, Ny+2, Nz+2] is the strided window starting at (plane 1, row 1, col 3).
 * x is the flux direction (reference dim 1): planes are the slab-partition and march axis.
 */
#ifndef TAUB200_H
#define TAUB200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define TAUB_ABI_VERSION 13
#define TAUB_GHOST 2            /* ghost width in x (planes), y (rows) and z (columns) */
#define TAUB_COL0 4             /* column of interior voxel k = 0 */
#define TAUB_MAX_LABELS 64      /* dense phase indices 0..L-1, L <= 64; index L = "outside" */

typedef enum taub_status {
    TAUB_OK = 0,
    TAUB_ERR_ARG = -1,          /* bad argument / unsupported shape */
    TAUB_ERR_CUDA = -2,         /* a CUDA call failed (message has the cudaError string) */
    TAUB_ERR_UNSUPPORTED = -3   /* path not available for this problem (e.g. fused sweep) */
} taub_status;

typedef enum taub_kind {
    TAUB_BINARY = 0,       /* Solver, PeriodicSolver */
    TAUB_MULTIPHASE = 1,   /* MultiPhaseSolver, PeriodicMultiPhaseSolver */
    TAUB_ANISOTROPIC = 2,  /* AnisotropicSolver (taufactor.py:422-478) through prefactor classes: codes = one uint16
                            * class id (< 64) per storage voxel, lut = device float[130]: {b, RN(1/b)} per class
                            * (b = weighted neighbour count :462-467; b = 1/b = 0 where it is infinite), then Ky, Kz */
    TAUB_MULTIPHASE_CLASS = 3  /* multi-phase through a stencil-class table: codes = one uint16 class id per
                                * storage voxel, lut = device float[L][8]: one 32-byte row
                                * {w_x+, w_x-, w_y+, w_y-, w_z+, w_z-, b, RN(1/b)} per class with b = prefactor
                                * (b = 1/b = 0 where it is infinite), most frequent classes first (the fused
                                * kernel keeps the first 256 rows in shared memory), L = number of classes
                                * (see taub_multiphase_keys) */
} taub_kind;

/* Geometry of one rank's slab.  Filled by taub_geom_init. */
typedef struct taub_geom {
    int32_t bs;                 /* images in the batch */
    int32_t Nx, Ny, Nz;         /* LOCAL voxel extents (Nx = planes owned by this rank) */
    int32_t Nx_global;          /* x extent of the whole volume */
    int32_t i_offset;           /* global x index of local plane 0 */
    int32_t periodic;           /* 1: y/z periodic (PeriodicSolver family) */
    int32_t planes, rows, pitch;/* storage extents: Nx+2G, Ny+2G, round_up(Nz+8, 32) */
    int64_t plane_stride;       /* rows * pitch   (elements) */
    int64_t image_stride;       /* planes * plane_stride */
} taub_geom;

/* Everything a sweep needs.  Mirrors the state SORSolver.__init__ builds (taufactor.py:24-67):
 * field (two ping-pong copies), the per-voxel prefactor in compressed form (4-bit neighbour
 * count per voxel for the binary solvers instead of the fp32 `factor` tensor (0 = non-conductive,
 * 1..8 = count, 9 = conductive voxel without a conductive neighbour); 1-byte dense
 * phase index + an (L+1)x(L+1) harmonic-mean table for the multi-phase solvers instead of
 * D_x, D_y, D_z, factor), and omega.  No chequerboard tensors: parity is computed on the fly. */
typedef struct taub_problem {
    taub_geom g;
    int32_t kind;               /* taub_kind */
    int32_t L;                  /* multi-phase: number of dense phase indices */
    float *field[2];            /* ping-pong buffers, bs * image_stride floats each */
    uint16_t *codes;            /* binary: bs * planes * rows * pitch/4 nibble-quads */
    uint8_t *labels;            /* multi-phase: bs * image_stride bytes */
    float *lut;                 /* multi-phase: (L+1)*(L+1) fp32 harmonic means, device */
    float omega;                /* over-relaxation factor rounded once to fp32 (taufactor.py:224) */
    int32_t cur;                /* index of the buffer that holds the current field */
    int32_t *stop;              /* optional device flag: while *stop != 0 every sweep / refresh / check
                                 * kernel of this problem returns at once (set by taub_check_async) */
    float *peer_lo[2];          /* optional, x-slab runs with equal slabs: the two ping-pong buffers of the rank */
    float *peer_hi[2];          /* below / above, mapped into this process (NVLink peer memory).  The fused
                                 * kernel then stores its first / last TAUB_GHOST output planes straight into
                                 * the neighbour's ghost planes of the destination buffer (one-sided halo
                                 * exchange; the caller orders passes with device-side signals). */
    int32_t *sync_ws;           /* optional: taub_sync_ws_ints() device int32 counters, zeroed once by the caller, for
                                 * the shared-memory resident path of small volumes (taub_resident_pairs); NULL =
                                 * that path is never taken */
    int32_t sync_epoch;         /* host-side: pairs run on sync_ws so far (maintained by taub_resident_pairs) */
    int32_t *redo_ws;           /* optional: taub_redo_ws_ints() device int32, zeroed once by the caller: lists of the
                                 * chunks a fused pass must redo with IEEE division (see taub_inexact_events); NULL =
                                 * such chunks are only counted */
} taub_problem;

/* -- library ------------------------------------------------------------------------------ */
int taub_abi_version(void);
const char *taub_last_error(void);
/* Device / toolchain facts for logs: SM count, compute capability, runtime version. */
int taub_device_info(int *sm_count, int *cc_major, int *cc_minor, int *runtime_version);
/* Number of kernels this library has launched since it was loaded (all threads). */
unsigned long long taub_launch_count(void);
/* Select the CUDA device for the calling thread (the library links its own static CUDA runtime,
 * whose current device is independent of PyTorch's).  Call before any other entry point. */
int taub_set_device(int ordinal);

/* -- geometry ----------------------------------------------------------------------------- */
int taub_geom_init(taub_geom *g, int bs, int Nx_local, int Ny, int Nz, int Nx_global,
                   int i_offset, int periodic);
size_t taub_field_elems(const taub_geom *g);   /* floats per ping-pong buffer */
size_t taub_codes_elems(const taub_geom *g);   /* uint16 elements */
size_t taub_sums_ws_bytes(const taub_geom *g); /* workspace for taub_plane_means */

/* -- construction (replaces taufactor.py:40-59: mask, vol_x, init_field :282-291,
 *    init_conductive_neighbours :402-410 / :493-499 / :585-604 / :626-650, _init_chequerboard) -- */
/* img: device uint8 [bs][img_n][Ny][Nz] holding global planes [img_i0, img_i0+img_n), which must
 * cover [i_offset-3, i_offset+Nx+3) clipped to the volume.  vec: device fp32 [Nx_global], the
 * reference's torch.linspace profile.  Fills both field buffers and the neighbour codes. */
int taub_init_binary(const taub_problem *p, const uint8_t *img, int img_i0, int img_n,
                     const float *vec, void *stream);
/* map256: device uint8[256] raw label -> dense index; cond: device fp32 [L+1], 1.0 where the
 * phase conducts (D > 0); p->lut must already hold the harmonic-mean table. */
int taub_init_multiphase(const taub_problem *p, const uint8_t *img, int img_i0, int img_n,
                         const uint8_t *map256, const float *cond, const float *vec, void *stream);
/* Multi-phase only: keys[b][i][j][k] (int32, device, bs*(i_hi-i_lo)*Ny*Nz) for the local planes [i_lo, i_hi)
 * (a slab may include its first ghost plane on either side, where the fused kernel also applies colour A)
 * = the seven dense phase indices that
 * determine a voxel's stencil (own | x- <<4 | x+ <<8 | y- <<12 | y+ <<16 | z- <<20 | z+ <<24) plus bit 28 /
 * 29 = first / last global plane (the Dirichlet face counts twice, taufactor.py:601-602).  Needs L <= 15.
 * Voxels with equal keys have identical face conductances and prefactor, so the caller can replace the
 * labels by the index into the table of distinct keys (TAUB_MULTIPHASE_CLASS). */
int taub_multiphase_keys(const taub_problem *p, int i_lo, int i_hi, int32_t *keys, void *stream);
/* The same classification without a per-voxel key array (what MultiPhaseSolver uses): taub_class_count fills a
 * device hash table with the distinct stencil keys of planes [i_lo, i_hi) and their voxel counts -- workspace of
 * taub_class_ws_bytes() bytes: int32 header[4] = {distinct keys, overflow flag (more than 65534 keys), capacity, 0},
 * int32 keys[capacity] (-1 = empty slot), uint64 counts[capacity].  The caller ranks the keys (most frequent first),
 * builds the weight rows (taufactor.py:594-603 in fp32) and passes slot_class[capacity] (uint16, device: class id of
 * the key in each slot); taub_class_assign then writes the class id of EVERY storage voxel into classes (uint16 per
 * storage voxel): voxels of planes [i_lo, i_hi) and, for the periodic solvers, their images in the y/z ghost frame
 * get their class, everything else `inert`. */
size_t taub_class_ws_bytes(void);
int taub_class_count(const taub_problem *p, int i_lo, int i_hi, void *ws, void *stream);
int taub_class_assign(const taub_problem *p, int i_lo, int i_hi, const void *ws, const uint16_t *slot_class, int inert,
                      uint16_t *classes, void *stream);
/* AnisotropicSolver state (taufactor.py:436-471): start field as taub_init_binary + one prefactor-class id per
 * storage voxel in p->codes: (nx * 3 + ny) * 3 + nz with nx = conductive x neighbours (the Dirichlet planes count 2:
 * 0..4), ny, nz in 0..2; 63 = non-conductive / outside.  p->kind == TAUB_ANISOTROPIC. */
int taub_init_anisotropic(const taub_problem *p, const uint8_t *img, int img_i0, int img_n, const float *vec, void *stream);
/* ElectrodeSolver / PeriodicElectrodeSolver state (electrode.py:39-64, :136-150; taufactor.py:47-56) of a whole volume:
 * class id per storage voxel in p->codes = b * 112 + (cond_nn * 7 + reac_nn) * 2 + [x+ neighbour conducts] for the
 * voxels of phase cond_label (cond_nn counts the left Dirichlet plane twice, the right end is closed; reac_nn =
 * neighbours of phase reac_label), bs * 112 elsewhere; start field vec[i] on the conductive phase, 2 (= 2 * left_bc)
 * on the left ghost plane; reac_sums[b][i] (int64, device) = sum of reac_nn over the conductive voxels of plane i.
 * img: device uint8 [bs][Nx][Ny][Nz]; p->kind == TAUB_MULTIPHASE_CLASS; a label of -1 matches nothing. */
int taub_init_electrode(const taub_problem *p, const uint8_t *img, int cond_label, int reac_label, const float *vec,
                        int64_t *reac_sums, void *stream);
/* counts[b][i] (int64, device) = voxels of local plane i whose raw label has sel256[label] != 0
 * (numerator of vol_x, taufactor.py:42); hist[b][256] (int64, device, may be NULL) = label
 * histogram (numerators of VF, taufactor.py:564-567). */
int taub_plane_counts(const taub_geom *g, const uint8_t *img, int img_i0, int img_n,
                      const uint8_t *sel256, int64_t *counts, int64_t *hist, void *stream);

/* -- the hot loop (replaces taufactor.py:174-182) ------------------------------------------ */
/* apply_boundary_conditions (taufactor.py:501-505, :652-656) on storage planes [p_lo, p_hi) of
 * one buffer: y/z ghost frame (width G, corners included) := periodic image of the interior. */
int taub_refresh_ghosts(const taub_geom *g, float *field, int p_lo, int p_hi, void *stream);
/* ONE reference iteration (one colour) field[cur] -> field[cur^1] on local planes
 * [i_lo, i_hi); iter = the reference's self.iter before the increment (selects the colour).
 * Generic path: any shape, both kinds.  Does not flip p->cur. */
int taub_half_sweep(const taub_problem *p, int64_t iter, int i_lo, int i_hi, void *stream);
/* TWO consecutive reference iterations (iter, iter+1) fused in one pass over HBM: temporally
 * blocked, TMA-bulk staged shared-memory tiles, register-rotating plane march.  Binary kind.
 * Returns TAUB_ERR_UNSUPPORTED when the problem does not qualify (see taub_can_fuse). */
int taub_fused_sweep2(const taub_problem *p, int64_t iter, int i_lo, int i_hi, void *stream);
int taub_can_fuse(const taub_problem *p);   /* binary, stencil-class and anisotropic kinds with their side arrays bound;
                                             * periodic problems with odd Ny / Nz run the OP kernel variant */
/* Diagnostics, no CUDA call: the launch plan taub_fused_sweep2 uses for planes [i_lo, i_hi) on a device with
 * `resident_ctas` CTA slots (2 x SM count).  out = { loaded tile rows, output tile rows, output float4 groups per tile
 * row, tile rows, tile columns, planes per chunk, plane chunks, grid rows (>= chunks: the chunk numbering may pad),
 * perm_R, perm_S (chunk of grid row y = (y % perm_R) * perm_S + y / perm_R; perm_R = 1: y), cluster width (0: none),
 * elastic chunk model (0 / 1) }.  Only p->g, p->kind, p->L and the non-nullness of the pointers are looked at. */
int taub_fused_plan(const taub_problem *p, int i_lo, int i_hi, int resident_ctas, int32_t out[12]);
/* The fused kernel divides by the neighbour count / prefactor with an FMA-corrected reciprocal that equals the IEEE
 * quotient for s == 0 and every |s| >= 2^-100.  A CTA whose threads met a non-zero value below 2^-100 (where that
 * sequence may be one subnormal ulp off) puts its chunk on a list in p->redo_ws, and a second kernel launched behind
 * every fused pass redoes the listed chunks with IEEE division -- the source buffer of a pass is read-only, so the
 * re-run simply overwrites the first run's output -- hence the fused trajectory equals the generic kernel's and the
 * reference's for every finite input.  (Environment TAUB_EXACT_REDO=0 or p->redo_ws == NULL: count only.)  This returns
 * how many chunks were listed since the library was loaded: 0 in the through-transport solves; the electrode solvers
 * get there when a cluster cut off from the inlet decays towards 0.  Synchronises the device. */
unsigned long long taub_inexact_events(void);
size_t taub_redo_ws_ints(void);                  /* int32 elements p->redo_ws must hold */
/* n iterations starting at iter on the whole local slab (single-rank use): refreshes periodic
 * ghosts, picks fused pairs where possible, flips p->cur.  flags bit 0: force the generic path;
 * bit 1: launch the queued kernels (fused / generic sweeps, ghost refresh) with programmatic dependent launch
 * (cudaLaunchAttributeProgrammaticStreamSerialization) -- the next pass's launch and shared-memory
 * prologue overlap the tail of the previous one; its first read waits for the previous grid to
 * complete (griddepcontrol.wait), so results are unchanged.  bit 2 (with bit 1, periodic solvers): the ghost
 * refresh releases the sweep behind it when it has finished, not when it starts.  bit 3: never take the
 * shared-memory resident path (taub_resident_pairs). */
int taub_iterate(taub_problem *p, int64_t iter, int n, int flags, void *stream);
/* Small whole volumes (binary kind, Nz >= 8, periodic only with even Ny and Nz, bricks of <= 227 KB): n_pairs x two
 * reference iterations (iter, iter+1, ...) in ONE cooperative launch with the field resident in shared memory --
 * one (x, y) brick per CTA, neighbours exchange their 2-wide faces through the ping-pong buffers with
 * release / acquire counters in p->sync_ws (no grid barrier, no relaunch, no pipeline fill per pass).  Same
 * arithmetic, bit-identical field.  Flips p->cur when n_pairs is odd.  taub_iterate takes this path by itself
 * when taub_can_reside(p) == 1 (flags bit 3 or TAUB_RESIDENT=0: never). */
int taub_can_reside(const taub_problem *p);
int taub_resident_pairs(taub_problem *p, int64_t iter, int n_pairs, void *stream);
size_t taub_sync_ws_ints(void);                 /* int32 counters p->sync_ws must hold */
/* Cycle counts of the phases of a pair, summed over the pairs that thread 0 of the middle brick has run since the last
 * reset: [0] wait for the neighbours' counters, [1] frame reload, [2] colour A, [3] colour B, [4] publish stores,
 * [5] fence + counter store, [6] number of pairs, [7] cycles of whole launches.  Synchronises the device. */
int taub_resident_profile(unsigned long long out[8], int reset);
unsigned long long taub_resident_timeouts(void); /* waits on a neighbour's counter that gave up after ~2 s (0 in
                                                  * every correct run; a non-zero value voids the result) */

/* -- the check (replaces vertical_flux :412-419 / :615-620 and the two torch.mean reductions in
 *    compute_metrics :296, :307) --------------------------------------------------------------- */
/* flux_mean[b][i], i in [0, Nx-1+has_next): mean over (y,z) of the (masked / weighted) flux
 * through the face between local planes i and i+1 (the last entry uses the upper ghost plane
 * and exists only when this slab is not the last: Nx_out = Nx-1 + (i_offset+Nx < Nx_global));
 * field_mean[b][i], i in [0, Nx): mean of the field over plane i.  fp64 accumulation in a fixed
 * order (deterministic), results rounded to fp32.  Outputs are device pointers.  While *p->stop != 0
 * (if p->stop is set) the outputs are left as they are. */
int taub_plane_means(const taub_problem *p, void *workspace, float *flux_mean, float *field_mean,
                     void *stream);

/* The same reduction followed by the reference's stop rule ON THE DEVICE (check_convergence,
 * taufactor.py:143-153, with compute_metrics :296-305 in the reference's float32 NumPy arithmetic,
 * including NumPy's pairwise summation order for the mean flux), so that the host can queue the next
 * block of sweeps without waiting for this check.  Needs the whole volume (i_offset == 0, Nx == Nx_global).
 *   D_mean[bs] device fp64; old_tau[bs] (in/out, starts at 0) and record[2 + 2*bs] (out) device fp32;
 *   record[0] = status (0 continue, 1 converged, 2 a slice flux is exactly 0 -> the host must run the
 *   percolation check, 3 the check did not run because *stop was already set), record[1] = reserved,
 *   record[2..2+bs) = tau, record[2+bs..2+2bs) = relative error.
 * On status 1 or 2 the kernel sets *p->stop (p->stop must be non-NULL): sweeps already queued behind
 * it become no-ops, so the field stays exactly at the iteration of this check. */
int taub_check_async(const taub_problem *p, void *workspace, float *flux_mean, float *field_mean,
                     const double *D_mean, float *old_tau, float conv_crit, float *record, void *stream);

/* x-slab runs: the stop rule alone, on per-slice flux means of the WHOLE volume that the caller has
 * assembled on the device (taub_plane_means of every slab, honouring p->stop, + an all-gather).  Same
 * arithmetic, record layout and stop-flag protocol as taub_check_async; every rank evaluates the same
 * numbers and therefore takes the same decision without a host round trip.
 *   flux_mean[bs][Nx_global - 1] device fp32; stop: this rank's device flag (non-NULL). */
int taub_stop_rule_async(int bs, int Nx_global, const float *flux_mean, const double *D_mean, float *old_tau,
                         float conv_crit, float *record, int32_t *stop, void *stream);


/* -- percolation check (replaces the host labelling of taufactor.py:318-327 ->
 *    metrics/connectivity.py:138-213, extract_through_feature(mask, 1, 'x')) ------------------- */
/* One round of a flood fill over the dense byte arrays mask / reach [bs][Nx][Ny][Nz] (device; mask != 0 =
 * conductive): forward + backward marches along x, y and z that carry reach through 6-connected mask
 * voxels (non-periodic).  The caller seeds reach with the mask of plane 0, repeats rounds until *changed
 * (device int, written by the call) stays 0, and then reads plane Nx-1 of reach: no reached voxel there
 * = no percolating path. */
int taub_flood_round(const uint8_t *mask, uint8_t *reach, int bs, int Nx, int Ny, int Nz, int32_t *changed,
                     void *stream);

/* ---- host-side helpers of taufactor_b200.io.imread (no CUDA): decoders of the two byte-oriented TIFF codecs
 * (TIFF 6.0 sections 9 and 13).  The reference's users read their volumes with tifffile.imread (README.md:51-54,
 * every notebook); both return the number of bytes written into dst (decoding stops when dst is full or src is
 * exhausted) or a negative status code for a corrupt LZW stream, message in taub_last_error(). */
int64_t taub_unpackbits(const uint8_t *src, size_t n_src, uint8_t *dst, size_t n_dst);
int64_t taub_unlzw(const uint8_t *src, size_t n_src, uint8_t *dst, size_t n_dst);

#ifdef __cplusplus
}
#endif
#endif /* TAUB200_H */
