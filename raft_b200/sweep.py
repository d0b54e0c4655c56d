"""Design sweeps sharded over GPUs (pattern of the reference's parametersweep.py:29-95).

(design) units are independent, so a sweep is an embarrassingly parallel shard: one process per GPU
(``torchrun``), rank r owns a contiguous block of designs, runs the fused solver on it, and the RAO blocks are
exchanged with ONE ``all_gather_into_tensor`` at the end -- the only collective on the path.  The helpers are
backend-agnostic (NCCL on GPUs; gloo on CPU tensors in the unit tests of the sharding logic).
"""
import copy

import numpy as np

from . import grid


def shard_bounds(n_items, rank, world):
    """Contiguous, balanced [lo, hi) of rank ``rank``: the first ``n_items % world`` ranks get one extra item."""
    base, extra = divmod(int(n_items), int(world))
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def all_gather_blocks(local, n_items, group=None):
    """All-gather per-rank blocks [n_local, ...] of a sharded array into [n_items, ...] (order = design index).

    Shards may be ragged by one item; every rank pads to the largest shard so a single
    ``all_gather_into_tensor`` suffices, then the padding is dropped."""
    import torch
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return local
    world = dist.get_world_size(group)
    n_max = -(-int(n_items) // world)
    pad = n_max - local.shape[0]
    buf = local if pad == 0 else torch.cat([local, local.new_zeros((pad,) + tuple(local.shape[1:]))], dim=0)
    out = local.new_empty((world * n_max,) + tuple(local.shape[1:]))
    dist.all_gather_into_tensor(out, buf.contiguous(), group=group)
    if n_max * world == n_items:
        return out
    keep = []
    for r in range(world):
        lo, hi = shard_bounds(n_items, r, world)
        keep.append(out[r * n_max:r * n_max + (hi - lo)])
    return torch.cat(keep, dim=0)


# ---- synthetic VolturnUS-S geometry variants (BASELINE.json configs[3]) ---------------------------------------

PARAMS = ("center_column_d", "outer_column_d", "draft", "outer_column_radius", "pontoon_height")


def sample_factors(n, seed=40, lower=0.75, upper=1.25):
    """[n,5] multiplicative factors ~ U[lower, upper] on the five parameters of parametersweep.py:29-44."""
    return np.random.default_rng(seed).uniform(lower, upper, size=(n, len(PARAMS)))


def apply_factors(base_design, f):
    """VolturnUS-S-like platform (center column, 3 outer columns, 3 pontoons, [upper supports]) with scaled
    centre-column diameter, outer-column diameter, draft, outer-column radius and pontoon height."""
    d = copy.deepcopy(base_design)
    mem = {m["name"]: m for m in d["platform"]["members"]}
    cc, oc, po = mem["center_column"], mem["outer_column"], mem["pontoon"]
    ccD, ocD = float(np.atleast_1d(cc["d"])[0]) * f[0], float(np.atleast_1d(oc["d"])[0]) * f[1]
    T = float(cc["rA"][2]) * f[2]
    ocR = float(oc["rA"][0]) * f[3]
    pH = float(po["d"][1]) * f[4]
    cc["d"], oc["d"] = ccD, ocD
    cc["rA"] = [cc["rA"][0], cc["rA"][1], T]
    oc["rA"] = [ocR, oc["rA"][1], T]
    oc["rB"] = [ocR, oc["rB"][1], oc["rB"][2]]
    po["d"] = [po["d"][0], pH]
    zp = T + pH / 2
    po["rA"] = [ccD / 2, po["rA"][1], zp]
    po["rB"] = [ocR - ocD / 2, po["rB"][1], zp]
    if "upper_support" in mem:
        us = mem["upper_support"]
        us["rA"] = [ccD / 2, us["rA"][1], us["rA"][2]]
        us["rB"] = [ocR - ocD / 2, us["rB"][1], us["rB"][2]]
    return d


def build_variants(base_design, base_matrices, factors, nw, max_freq, depth):
    """Packed designs (``packer.pack_fowt`` dicts) for every row of ``factors`` on a grid of nw bins.

    Node tables and the Morison added mass follow the geometry (``raft_b200.member``); structural mass,
    hydrostatic and mooring stiffness are statics (out of scope) and stay at ``base_matrices``."""
    from .fowt import FOWT
    w = grid.make_w(max_freq / nw, max_freq)
    k = grid.wave_number(w, depth)
    out = []
    for f in np.atleast_2d(factors):
        fw = FOWT(apply_factors(base_design, f), w, depth=depth, matrices=base_matrices, k=k)
        fw.calcHydroConstants()
        out.append(fw.pack())
    return out


def solve_sweep(packed_designs, cases, n_iter=10, tol=0.01, xi_start=0.0, device=None, group=None, n_total=None):
    """Solve this rank's designs on its GPU and all-gather the RAOs: -> (Xi [n_total,nC,6,nw], status [n_total,nC,4])."""
    from . import solver
    sess = solver.DeviceSession(solver.DesignBatch(packed_designs), solver.CaseTable(cases), device=device)
    out = sess.solve(n_iter=n_iter, tol=tol, xi_start=xi_start)
    n_total = len(packed_designs) if n_total is None else n_total
    return all_gather_blocks(out["Xi"], n_total, group), all_gather_blocks(out["status"], n_total, group)


class PipelinedSolve:
    """Solve this rank's units in ``n_chunks`` launches and overlap each chunk's all-gather (NCCL, side stream)
    with the next chunk's kernels -- the transfer of chunk i hides behind the compute of chunk i+1.

    ``split="designs"`` chunks the design list (sweeps), ``split="cases"`` the case table (one design, many sea
    states).  ``gathered[i]`` is [world, ...] for chunk i; with one rank nothing is gathered."""

    def __init__(self, packed_designs, cases, n_chunks=2, split="designs", device=None, group=None, want=("Xi", "status")):
        import torch
        import torch.distributed as dist
        from . import solver
        self.torch, self.dist, self.group = torch, dist, group
        self.world = dist.get_world_size(group) if (dist.is_available() and dist.is_initialized()) else 1
        designs = [packed_designs] if isinstance(packed_designs, dict) else list(packed_designs)
        n = len(designs) if split == "designs" else len(cases["Hs"])
        n_chunks = max(1, min(n_chunks, n))
        self.sessions = []
        for i in range(n_chunks):
            lo, hi = shard_bounds(n, i, n_chunks)
            if split == "designs":
                sess = solver.DeviceSession(solver.DesignBatch(designs[lo:hi]), solver.CaseTable(cases), device=device, want=want)
            else:
                sub = {k: np.asarray(v)[lo:hi] for k, v in cases.items()}
                sess = solver.DeviceSession(solver.DesignBatch(designs), solver.CaseTable(sub), device=device, want=want)
            self.sessions.append(sess)
        dev = self.sessions[0].device
        self.comm = torch.cuda.Stream(device=dev) if self.world > 1 else None
        self.gathered = [torch.empty((self.world,) + tuple(s.out["Xi"].shape), dtype=s.out["Xi"].dtype, device=dev)
                         if self.world > 1 else None for s in self.sessions]
        self._ag_done = [None] * len(self.sessions)
        self.units = sum(s.batch.n_designs * s.cases.n_cases * s.batch.nw for s in self.sessions)

    def step(self, **solve_kw):
        """Enqueue one step.  Software pipeline: chunk i's kernels only wait for the all-gather that last read chunk
        i's output buffer (issued one step ago), so gathers also overlap the NEXT step's kernels; call ``drain()``
        before reading ``gathered`` or stopping a timer."""
        torch = self.torch
        cur = torch.cuda.current_stream(self.sessions[0].device)
        for i, (sess, g) in enumerate(zip(self.sessions, self.gathered)):
            if self.world > 1 and self._ag_done[i] is not None:
                cur.wait_event(self._ag_done[i])            # the previous gather of this buffer has consumed it
            sess.solve(**solve_kw)
            if self.world > 1:
                ev = torch.cuda.Event()
                ev.record(cur)
                with torch.cuda.stream(self.comm):
                    self.comm.wait_event(ev)
                    self.dist.all_gather_into_tensor(g, sess.out["Xi"], group=self.group)
                    done = torch.cuda.Event()
                    done.record(self.comm)
                self._ag_done[i] = done

    def drain(self):
        """Make the current stream wait for every outstanding all-gather."""
        if self.world > 1:
            self.torch.cuda.current_stream(self.sessions[0].device).wait_stream(self.comm)

    def status(self):
        return np.concatenate([s.out["status"].cpu().numpy().reshape(-1, 4) for s in self.sessions], axis=0)
