"""TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).

PyTorch-eager CPU restatement of the reference's hot loop, executed the way the reference executes
it on a CPU: one library elementwise kernel per arithmetic step with full-size temporaries
(ref:174-182, :95-103, :606-613).  It exists so bench.py can time "the reference's CPU path" on
the GPU box, where /root/reference does not exist; it reuses the NumPy oracle's state builder and
is checked bit-for-bit against it in tests/test_oracle_golden.py."""
import numpy as np
import torch


def from_state(st):
    """Torch mirrors of a sor_numpy state (fp32, CPU)."""
    t = dict(kind=st["kind"], periodic=st["periodic"], iter=st["iter"],
             field=torch.from_numpy(np.array(st["field"], dtype=np.float32, copy=True)),
             factor=torch.from_numpy(np.ascontiguousarray(st["factor"])),
             cb=[torch.from_numpy(np.ascontiguousarray(c)) for c in st["cb"]])
    if st["kind"] == "multiphase":
        for k in ("D_x", "D_y", "D_z"):
            t[k] = torch.from_numpy(np.ascontiguousarray(st[k]))
    return t


def half_sweep(t):
    """One reference iteration, same op sequence as ref:175-182."""
    f = t["field"]
    with torch.no_grad():
        if t["periodic"]:
            f[:, :, 0, :] = f[:, :, -2, :]
            f[:, :, -1, :] = f[:, :, 1, :]
            f[:, :, :, 0] = f[:, :, :, -2]
            f[:, :, :, -1] = f[:, :, :, 1]
        if t["kind"] == "binary":
            inc = f[:, 2:, 1:-1, 1:-1] + f[:, :-2, 1:-1, 1:-1] + f[:, 1:-1, 2:, 1:-1] + \
                f[:, 1:-1, :-2, 1:-1] + f[:, 1:-1, 1:-1, 2:] + f[:, 1:-1, 1:-1, :-2]
        else:
            Dx, Dy, Dz = t["D_x"], t["D_y"], t["D_z"]
            inc = f[:, 2:, 1:-1, 1:-1] * Dx[:, 1:] + f[:, :-2, 1:-1, 1:-1] * Dx[:, :-1] + \
                f[:, 1:-1, 2:, 1:-1] * Dy[:, :, 1:] + f[:, 1:-1, :-2, 1:-1] * Dy[:, :, :-1] + \
                f[:, 1:-1, 1:-1, 2:] * Dz[:, :, :, 1:] + f[:, 1:-1, 1:-1, :-2] * Dz[:, :, :, :-1]
        inc /= t["factor"]
        inc -= f[:, 1:-1, 1:-1, 1:-1]
        inc *= t["cb"][t["iter"] % 2]
        f[:, 1:-1, 1:-1, 1:-1] += inc
    t["iter"] += 1


def flux_check(t):
    """The per-check reductions of ref:293-307 (binary / multi-phase), returned as numpy."""
    f = t["field"]
    with torch.no_grad():
        vf = f[:, 2:-1, 1:-1, 1:-1] - f[:, 1:-2, 1:-1, 1:-1]
        if t["kind"] == "binary":
            vf[t["factor"][:, 0:-1] > 8] = 0
            vf[t["factor"][:, 1:] > 8] = 0
        else:
            vf = t["D_x"][:, 1:-1] * vf
        return torch.mean(vf, (2, 3)).numpy(), torch.mean(f[:, 1:-1, 1:-1, 1:-1], (2, 3)).numpy()
