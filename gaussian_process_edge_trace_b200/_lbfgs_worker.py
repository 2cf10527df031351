"""Host side of the final hyper-parameter fit: many independent L-BFGS-B minimisations advanced in lock step.

Every instance is driven by scipy's own reverse-communication routine `scipy.optimize._lbfgsb.setulb` with the
constants `scipy.optimize.minimize(method='L-BFGS-B')` uses (what the reference calls at sklearn_gpr.py:589-595), so
the iterates are the ones scipy would produce for the same objective values; only the objective/gradient
evaluations are handed out (to the GPU, in batches). This module imports numpy/scipy only, so worker processes
started with `spawn` stay light.
"""
import numpy as np
from scipy.optimize import _lbfgsb

M_CORR = 10                       # maxcor
FACTR = 2.220446049250313e-09 / np.finfo(float).eps   # ftol / eps
PGTOL = 1e-5
MAXLS = 20
MAXITER = 15000
MAXFUN = 15000
try:
    from scipy.optimize._lbfgsb_py import HAS_ILP64
except Exception:  # pragma: no cover
    HAS_ILP64 = False
_INT = np.int64 if HAS_ILP64 else np.int32


class Instance:
    """One L-BFGS-B run (state arrays exactly as scipy's _minimize_lbfgsb allocates them)."""
    __slots__ = ("x", "f", "g", "low", "up", "nbd", "wa", "iwa", "task", "ln_task", "lsave", "isave", "dsave", "nit",
                 "nfev")

    def __init__(self, x0, low, up):
        n = x0.shape[0]
        m = M_CORR
        self.x = np.array(np.clip(x0, low, up), dtype=np.float64)
        self.f = np.array(0.0, dtype=np.float64)
        self.g = np.zeros((n,), dtype=np.float64)
        self.low = np.array(low, dtype=np.float64)
        self.up = np.array(up, dtype=np.float64)
        self.nbd = np.full(n, 2, dtype=_INT)          # both bounds finite
        self.wa = np.zeros(2 * m * n + 5 * n + 11 * m * m + 8 * m, np.float64)
        self.iwa = np.zeros(3 * n, dtype=_INT)
        self.task = np.zeros(2, dtype=_INT)
        self.ln_task = np.zeros(2, dtype=_INT)
        self.lsave = np.zeros(4, dtype=_INT)
        self.isave = np.zeros(44, dtype=_INT)
        self.dsave = np.zeros(29, dtype=np.float64)
        self.nit = 0
        self.nfev = 0

    def advance(self):
        """Runs setulb until it asks for f, g at self.x (returns True) or stops (returns False)."""
        while True:
            _lbfgsb.setulb(M_CORR, self.x, self.low, self.up, self.nbd, self.f, self.g, FACTR, PGTOL, self.wa, self.iwa,
                           self.task, self.lsave, self.isave, self.dsave, MAXLS, self.ln_task)
            t = self.task[0]
            if t == 3:
                return True
            if t == 1:
                self.nit += 1
                if self.nit >= MAXITER:
                    self.task[0], self.task[1] = 5, 504
                elif self.nfev > MAXFUN:
                    self.task[0], self.task[1] = 5, 502
            else:
                return False

    def give(self, f, g):
        self.f = np.array(f, dtype=np.float64)
        self.g = np.asarray(g, dtype=np.float64).copy()
        self.nfev += 1


class Shard:
    """A set of instances advanced together. `begin` returns the points whose objective is needed; `feed` takes the
    values for exactly those points (same order) and returns the next set."""

    def __init__(self):
        self.inst, self.pending = [], []

    def begin(self, x0, low, up):
        self.inst = [Instance(x0[i], low, up) for i in range(x0.shape[0])]
        self.pending = [i for i, s in enumerate(self.inst) if s.advance()]
        return self._points()

    def _points(self):
        idx = np.array(self.pending, dtype=np.int64)
        pts = np.stack([self.inst[i].x for i in self.pending]) if self.pending else np.zeros((0, 3))
        return idx, pts

    def feed(self, f, g):
        nxt = []
        for k, i in enumerate(self.pending):
            s = self.inst[i]
            s.give(f[k], g[k])
            if s.advance():
                nxt.append(i)
        self.pending = nxt
        return self._points()

    def result(self):
        return (np.stack([s.x for s in self.inst]), np.array([float(s.f) for s in self.inst]),
                np.array([s.nfev for s in self.inst]))


def worker_main(rd, wr):
    """Worker loop over a pair of binary streams (pickle framing)."""
    import pickle
    shard = Shard()
    while True:
        try:
            msg = pickle.load(rd)
        except EOFError:
            return
        op = msg[0]
        if op == "begin":
            out = shard.begin(msg[1], msg[2], msg[3])
        elif op == "feed":
            out = shard.feed(msg[1], msg[2])
        elif op == "result":
            out = shard.result()
        elif op == "stop":
            return
        else:
            out = None
        pickle.dump(out, wr, protocol=pickle.HIGHEST_PROTOCOL)
        wr.flush()


class _Conn:
    """Parent end of one worker: a plain subprocess (`python -m ..._lbfgs_worker`), independent of the parent's
    __main__ module, start method and CUDA state."""

    def __init__(self):
        import os
        import subprocess
        import sys
        root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
        env = dict(os.environ)
        env["PYTHONPATH"] = root + os.pathsep + env.get("PYTHONPATH", "")
        env["OPENBLAS_NUM_THREADS"] = env["OMP_NUM_THREADS"] = "1"
        self.proc = subprocess.Popen([sys.executable, "-m", "gaussian_process_edge_trace_b200._lbfgs_worker"],
                                     stdin=subprocess.PIPE, stdout=subprocess.PIPE, env=env)

    def send(self, obj):
        import pickle
        pickle.dump(obj, self.proc.stdin, protocol=pickle.HIGHEST_PROTOCOL)
        self.proc.stdin.flush()

    def recv(self):
        import pickle
        return pickle.load(self.proc.stdout)

    def close(self):
        try:
            self.send(("stop",))
            self.proc.stdin.close()
        except Exception:
            pass
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()


class LbfgsbPool:
    """n_workers processes (0 = in-process) that hold the instances; the caller supplies `evaluate(ids, thetas)`."""

    def __init__(self, n_workers):
        self.n_workers = n_workers
        self.local = Shard() if n_workers == 0 else None
        self.conns = [_Conn() for _ in range(n_workers)]

    def close(self):
        for c in self.conns:
            c.close()
        self.conns = []

    def minimize_many(self, x0, low, up, submit, wait, n_groups=2):
        """x0[E, n]. `submit(group, ids int64[k], thetas float64[k, n])` starts an (asynchronous) evaluation and
        returns a handle; `wait(handle)` -> (f[k], g[k, n]). The workers are split into `n_groups` groups that are
        evaluated alternately, so the GPU evaluates one group while the other group's workers advance.
        Returns (x[E, n], f[E], nfev[E], rounds)."""
        E = x0.shape[0]
        if self.n_workers == 0:
            idx, pts = self.local.begin(x0, low, up)
            rounds = 0
            while idx.shape[0]:
                f, g = wait(submit(0, idx, pts))
                rounds += 1
                idx, pts = self.local.feed(f, g)
            xs, fs, nf = self.local.result()
            return xs, fs, nf, rounds
        W = min(self.n_workers, E)
        cuts = np.linspace(0, E, W + 1).astype(int)
        parts = [(int(cuts[k]), int(cuts[k + 1])) for k in range(W)]
        G = max(1, min(n_groups, W))
        groups = [list(range(gi, W, G)) for gi in range(G)]
        for k in range(W):
            self.conns[k].send(("begin", x0[parts[k][0]:parts[k][1]], low, up))
        replies = {k: self.conns[k].recv() for k in range(W)}
        rounds = 0

        def launch(gi):
            ks = [k for k in groups[gi] if replies[k][0].shape[0]]
            if not ks:
                return None
            ids = np.concatenate([replies[k][0] + parts[k][0] for k in ks])
            pts = np.concatenate([replies[k][1] for k in ks], axis=0)
            return ks, [replies[k][0].shape[0] for k in ks], submit(gi, ids, pts)

        inflight = [launch(gi) for gi in range(G)]
        while any(h is not None for h in inflight):
            for gi in range(G):
                h = inflight[gi]
                if h is None:
                    continue
                ks, counts, handle = h
                f, g = wait(handle)
                rounds += 1
                offs = np.concatenate([[0], np.cumsum(counts)])
                for j, k in enumerate(ks):
                    self.conns[k].send(("feed", f[offs[j]:offs[j + 1]], g[offs[j]:offs[j + 1]]))
                # while these workers advance, the other groups' evaluations are (or get) in flight
                inflight[gi] = ("recv", ks)
            for gi in range(G):
                h = inflight[gi]
                if h is None:
                    continue
                for k in h[1]:
                    replies[k] = self.conns[k].recv()
                inflight[gi] = launch(gi)
        for k in range(W):
            self.conns[k].send(("result",))
        outs = [self.conns[k].recv() for k in range(W)]
        xs = np.concatenate([o[0] for o in outs])
        fs = np.concatenate([o[1] for o in outs])
        nf = np.concatenate([o[2] for o in outs])
        return xs, fs, nf, rounds


if __name__ == "__main__":
    import sys
    worker_main(sys.stdin.buffer, sys.stdout.buffer)
