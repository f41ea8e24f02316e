"""Seeded synthetic inputs and the parity-case catalogue shared by the golden generator
(tests/golden/make_golden.py, runs the real reference in the build container), the oracle tests
(CPU) and the CUDA parity tests (GPU).  Generators follow SURVEY.md section 8(d); the structured
images restate the inputs of /root/reference/tests/test_taufactor.py (cited per case)."""
import numpy as np


# ----------------------------------------------------------------------------- generators
def random_img(shape, p=0.5, seed=0):
    if isinstance(shape, int):
        shape = (shape,) * 3
    return (np.random.default_rng(seed).random(shape) < p).astype(np.uint8)


def _smooth_noise(shape, blobiness, seed):
    from scipy.ndimage import gaussian_filter
    g = np.random.default_rng(seed).random(shape, dtype=np.float32)
    return gaussian_filter(g, sigma=float(np.mean(shape)) / (40.0 * blobiness), mode="wrap")


def blobs(shape, porosity=0.5, blobiness=1, seed=0):
    """Binary periodic blob structure, 1 = conductive, volume fraction = porosity."""
    if isinstance(shape, int):
        shape = (shape,) * 3
    g = _smooth_noise(shape, blobiness, seed)
    return (g > np.quantile(g, 1 - porosity)).astype(np.uint8)


def blobs3(shape, fractions=(0.40, 0.35, 0.25), blobiness=1, seed=0):
    """Three-phase periodic blob structure with labels {0,1,2}."""
    if isinstance(shape, int):
        shape = (shape,) * 3
    g = _smooth_noise(shape, blobiness, seed)
    return np.digitize(g, np.quantile(g, np.cumsum(fractions)[:-1])).astype(np.uint8)


def generate(job):
    """(kind, size, seed) -> image; picklable entry point for worker processes that build big volumes."""
    kind, size, seed = job
    if kind == "blobs":
        return blobs(size, 0.5, seed=seed)
    if kind == "blobs3":
        return blobs3(size, seed=seed)
    raise ValueError(kind)


def _cache_path(job):
    import os
    import tempfile
    d = os.path.join(tempfile.gettempdir(), "taub_img_cache")
    os.makedirs(d, exist_ok=True)
    return os.path.join(d, "{}_{}_{}.npy".format(*job))


def generate_parallel(jobs, workers=None, cache=True):
    """Images of ``jobs`` = [(kind, size, seed), ...], built in parallel child interpreters (``python cases.py kind
    size seed out.npy``: plain subprocesses, so a CUDA context in the parent does not matter) and kept as .npy files
    under the system temp directory, so the two arms of bench.py and the tests of one box generate each volume once."""
    import os
    import subprocess
    import sys
    workers = workers or max(1, min(len(jobs), (os.cpu_count() or 2) - 1))
    out = [None] * len(jobs)
    todo, running = [], []
    for i, job in enumerate(jobs):
        path = _cache_path(job)
        if cache and os.path.exists(path):
            try:
                out[i] = np.load(path)
                continue
            except Exception:
                os.remove(path)
        todo.append((i, job, path))
    if len(todo) == 1 and not cache:
        i, job, _ = todo.pop()
        out[i] = generate(job)
    while todo or running:
        while todo and len(running) < workers:
            i, (kind, size, seed), path = todo.pop(0)
            tmp = path + f".{os.getpid()}.tmp.npy"
            running.append((i, path, tmp, subprocess.Popen([sys.executable, os.path.abspath(__file__), kind, str(size),
                                                            str(seed), tmp])))
        i, path, tmp, proc = running.pop(0)
        if proc.wait() != 0:
            raise RuntimeError(f"image generator failed for job {jobs[i]}")
        out[i] = np.load(tmp)
        if cache:
            os.replace(tmp, path)
        else:
            os.remove(tmp)
    return out


def uniform_block(shape, zero_row=True):
    img = np.ones(shape)
    if zero_row:
        img[:, 0] = 0
    return img


def head_only(N=20):
    img = np.zeros((N, N, N))
    img[:2] = 1
    return img


def strip(N=20, t=10):
    img = np.zeros((N, N, N))
    img[:, 0:t, 0:t] = 1
    return img


def slanted_strip(N=20):
    img = np.zeros((N, N + 1, N + 1))
    for i in range(N):
        img[i, i:i + 2, i:i + 2] = 1
    return img


def deadend():
    solid = np.zeros((10, 50, 50))
    solid[:8, 25, 25] = 1
    return solid


def strip12(N=20, x=10):
    img = np.zeros((N, N, N))
    img[:, 0:x, 0:x] = 1
    img[:, 0:x, x:N] = 2
    return img


def strip123(N=20, x=10):
    img = np.ones((N, N, N))
    img[:, 0:x, 0:x] = 2
    img[:, 0:x, x:N] = 3
    return img


def label0_conductive(N=20):
    img = np.zeros((N, N, N))
    img[:, :2] = 1
    return img


def batched_mp(N=16):
    a = np.ones((N, N, N))
    b = np.ones((N, N, N))
    b[:, :, : N // 2] = 2
    return np.stack([a, b], axis=0)


def stacked_blobs(n=3, N=48, seed0=384):
    return np.stack([blobs(N, 0.5, seed=seed0 + b) for b in range(n)])


def odd_random(seed=7):
    return random_img((11, 13, 9), 0.7, seed)


def odd_random3(seed=9):
    return (np.random.default_rng(seed).random((10, 9, 7)) * 3).astype(np.uint8)


def flat_2d(seed=3):
    return random_img((24, 31), 0.75, seed)  # 2-D input -> Nz = 1 (ref:198-199)


def img_2d_batch(seed=5):
    return (np.random.default_rng(seed).random((3, 30, 28, 1)) < 0.8).astype(np.uint8)


# ----------------------------------------------------------------------------- catalogue
# name -> (solver class name, image builder, ctor kwargs, solve kwargs, expectation from the
#          reference's own tests or None)
CASES = {
    # reference tests/test_taufactor.py:12-76  (Solver)
    "ref_uniform20":        ("Solver", lambda: uniform_block((20, 20, 20)), {}, {}, ("tau5", 1.0)),
    "ref_rect_solver_dim":  ("Solver", lambda: uniform_block((40, 20, 20)), {}, {}, ("tau5", 1.0)),
    "ref_rect_other_dim":   ("Solver", lambda: uniform_block((20, 20, 40)), {}, {}, ("tau5", 1.0)),
    "ref_non_percolating":  ("Solver", head_only, {}, {"iter_limit": 1000}, ("inf", None)),
    "ref_strip":            ("Solver", strip, {}, {}, ("tau5", 1.0)),
    "ref_slanted":          ("Solver", slanted_strip, {}, {}, ("tau5", 7.51667)),
    "ref_deadend":          ("Solver", deadend, {}, {}, ("inf", None)),
    # :80-106 (PeriodicSolver)
    "ref_per_uniform20":    ("PeriodicSolver", lambda: uniform_block((20, 20, 20)), {}, {}, ("tau5", 1.0)),
    "ref_per_non_perc":     ("PeriodicSolver", head_only, {}, {"iter_limit": 1000}, ("inf", None)),
    "ref_per_strip":        ("PeriodicSolver", strip, {}, {}, ("tau5", 1.0)),
    # :110-232 (MultiPhaseSolver)
    "ref_mp_non_perc":      ("MultiPhaseSolver", head_only, {}, {"iter_limit": 1000}, ("inf", None)),
    "ref_mp_ones":          ("MultiPhaseSolver", lambda: np.ones((20, 20, 20)), {}, {"iter_limit": 1000}, ("tau4", 1.0)),
    "ref_mp_halves":        ("MultiPhaseSolver", lambda: np.ones((20, 20, 20)), {"diffusivities": {1: 0.5}}, {"iter_limit": 1000}, ("tau4", 1.0)),
    "ref_mp_strip":         ("MultiPhaseSolver", strip, {}, {}, ("tau4", 1.0)),
    "ref_mp_strip12":       ("MultiPhaseSolver", strip12, {"diffusivities": {0: 0, 1: 1, 2: 0.5}}, {}, ("tau4", 1.0)),
    "ref_mp_strip123":      ("MultiPhaseSolver", strip123, {"diffusivities": {0: 0, 1: 1, 2: 0.5, 3: 2}}, {}, ("tau4", 1.0)),
    "ref_mp_label0":        ("MultiPhaseSolver", label0_conductive, {"diffusivities": {0: 1.0, 1: 0.0}}, {"iter_limit": 1000}, ("tau4", 1.0)),
    "ref_mp_slanted":       ("MultiPhaseSolver", slanted_strip, {}, {"iter_limit": 1000}, None),
    "ref_mp_batched":       ("MultiPhaseSolver", batched_mp, {"diffusivities": {0: 0.0, 1: 1.0, 2: 0.5}}, {"iter_limit": 1000}, None),
    "ref_pmp_uniform":      ("PeriodicMultiPhaseSolver", lambda: np.ones((20, 20, 20)), {"diffusivities": {1: 1.0}}, {"iter_limit": 1000}, ("tau4", 1.0)),
    "ref_per_slanted_odd":  ("PeriodicSolver", slanted_strip, {}, {"iter_limit": 1000}, None),
    "ref_pmp_slanted_odd":  ("PeriodicMultiPhaseSolver", slanted_strip, {"diffusivities": {0: 0.0, 1: 1.0}}, {"iter_limit": 1000}, None),
    # seeded synthetic cases
    "rand40":               ("Solver", lambda: random_img(40, 0.6, 0), {}, {}, None),
    "rand100_cfg1":         ("Solver", lambda: random_img(100, 0.5, 0), {}, {}, None),          # BASELINE config 1
    "blobs64":              ("Solver", lambda: blobs(64, 0.5, seed=64), {}, {}, None),
    "blobs64_per":          ("PeriodicSolver", lambda: blobs(64, 0.5, seed=64), {}, {}, None),
    "blobs3_48_mp":         ("MultiPhaseSolver", lambda: blobs3(48, seed=768), {"diffusivities": {0: 0.0, 1: 1.0, 2: 0.3}}, {}, None),
    "blobs3_48_pmp":        ("PeriodicMultiPhaseSolver", lambda: blobs3(48, seed=768), {"diffusivities": {0: 0.0, 1: 1.0, 2: 0.3}}, {}, None),
    "blobs3_96_mp":         ("MultiPhaseSolver", lambda: blobs3(96, seed=768), {"diffusivities": {0: 0.0, 1: 1.0, 2: 0.3}}, {}, None),
    "blobs3_96_pmp":        ("PeriodicMultiPhaseSolver", lambda: blobs3(96, seed=768), {"diffusivities": {0: 0.0, 1: 1.0, 2: 0.3}}, {}, None),
    "batch3_blobs48":       ("Solver", stacked_blobs, {}, {}, None),                           # joint stop rule
    "odd_11_13_9":          ("Solver", odd_random, {}, {"iter_limit": 300}, None),
    "odd_11_13_9_per":      ("PeriodicSolver", odd_random, {}, {"iter_limit": 300}, None),
    "odd3_mp":              ("MultiPhaseSolver", odd_random3, {"diffusivities": {0: 0.2, 1: 1.0, 2: 0.0}}, {"iter_limit": 300}, None),
    "odd3_pmp":             ("PeriodicMultiPhaseSolver", odd_random3, {"diffusivities": {0: 0.2, 1: 1.0, 2: 0.0}}, {"iter_limit": 300}, None),
    "flat2d":               ("Solver", flat_2d, {}, {"iter_limit": 600}, None),
    "flat2d_per_batch":     ("PeriodicSolver", img_2d_batch, {}, {"iter_limit": 600}, None),
    # AnisotropicSolver (ref:422-478; SURVEY 8f "next" #1)
    "aniso_iso":            ("AnisotropicSolver", lambda: random_img((24, 20, 16), 0.7, 11), {"spacing": (1, 1, 1)}, {}, None),
    "aniso_fib":            ("AnisotropicSolver", lambda: random_img((20, 24, 18), 0.75, 12), {"spacing": (1.0, 1.0, 2.5)}, {}, None),
    "aniso_odd":            ("AnisotropicSolver", lambda: odd_random(8), {"spacing": (2.0, 1.0, 3.0)}, {"iter_limit": 300}, None),
    "aniso_blobs48":        ("AnisotropicSolver", lambda: blobs(48, 0.5, seed=48), {"spacing": (1.0, 0.8, 1.6)}, {}, None),
    "omega_custom":         ("Solver", lambda: random_img((24, 20, 16), 0.7, 11), {"omega": 1.7}, {"conv_crit": 1e-3}, None),
}

# cases whose field is snapshotted bit-for-bit after these iteration counts
SNAPSHOT_ITERS = (1, 2, 3, 100, 101)
SNAPSHOT_CASES = ("odd_11_13_9", "odd_11_13_9_per", "odd3_mp", "odd3_pmp", "flat2d",
                  "flat2d_per_batch", "ref_slanted", "ref_per_slanted_odd", "aniso_odd", "aniso_fib")

# fast subset for the GPU parity run through the product API
GPU_SOLVE_CASES = tuple(CASES)


if __name__ == "__main__":      # worker of generate_parallel
    import sys
    np.save(sys.argv[4], generate((sys.argv[1], int(sys.argv[2]), int(sys.argv[3]))))
