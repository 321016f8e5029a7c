"""TEST INFRASTRUCTURE: an engine adapter for spral_b200.dist that puts the reference's own CPU
engine (oracle/_ref through tests/oracle_ref.py) in the place of the B200 engine, so that the host logic
of the one-process-per-GPU driver -- part ownership, contribution hand-over across ranks, the solve
exchanges, the final all-reduce -- runs in the world_size-2 gloo tests on CPU.  The product package
contains no reference to it (spral_b200.dist.GpuEngine is the only engine there)."""
import ctypes as C

import oracle_ref


class OracleEngine:
    device_ipc = False            # contribution blocks are host arrays: staged through the rendezvous store

    def device(self, local_rank):
        import torch
        return torch.device("cpu")

    def symbolic(self, analysis, part, local_rank, options):
        return ("oracle", part)   # the reference builds its symbolic subtree inside RefSubtree

    def factor(self, symb, analysis, part, posdef, val, child_contrib, options, scaling):
        return oracle_ref.RefSubtree(analysis, part, posdef, val, child_contrib, options, scaling)

    def solve(self, ns, which, X, nrhs, n):
        f = getattr(oracle_ref.load(), f"spral_ssids_cpu_subtree_solve_{which}_dbl")
        rc = f(ns.posdef, ns.h, nrhs, C.c_void_p(X.data_ptr()), n)
        assert rc == 0

    def get_contrib(self, ns):
        return ns.get_contrib()

    def device_ms(self, ns):
        return 0.0
