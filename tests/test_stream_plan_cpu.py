"""CPU: the host-side planner of the streaming schedule (bhmc_stream_plan_host, no device needed).

An interpreter replays the op tables phase by phase, exactly in the order the launch loop issues the kernels
(update-1, kinetic/accept, begin, update-2, gradient launch), and checks that every chain experiences the reference's
sequence of events (inference/cpu/hmc.py:39-64): momentum draw, (L-1) x [for each variable: half kick + drift,
gradient, full kick], Metropolis test -- with the gradient at a transition's start point taken from the previous
transition (never re-evaluated after the first), and that all rows move the same sweep group before any one launch.
"""
import ctypes as C

import numpy as np
import pytest

from dropout_hamiltonian_montecarlo_b200 import _lib

OP_LATCH, OP_POST, OP_PRE, OP_FINISH, OP_BEGIN, OP_LATCH_CACHED = 1, 2, 4, 8, 16, 32


def plan(L, nsw):
    L = np.ascontiguousarray(L, dtype=np.int32)
    n_steps, n_chains = L.shape
    lib = _lib.lib()
    J, ng = C.c_int64(), C.c_int64()
    p = lambda a: a.ctypes.data_as(C.c_void_p) if a is not None else None
    _lib.check(lib.bhmc_stream_plan_host(p(L), n_chains, n_steps, nsw, C.byref(J), C.byref(ng), None, None, None, None,
                                         None, None, None))
    n = (J.value + 1) * n_chains
    code1, code2 = np.zeros(n, np.uint32), np.zeros(n, np.uint32)
    step1, step2 = np.zeros(n, np.int32), np.zeros(n, np.int32)
    perm = np.zeros(n_chains, np.int32)
    rows_el, rows_grad = np.zeros(J.value + 1, np.int32), np.zeros(J.value + 1, np.int32)
    _lib.check(lib.bhmc_stream_plan_host(p(L), n_chains, n_steps, nsw, C.byref(J), C.byref(ng), p(code1), p(code2), p(step1),
                                         p(step2), p(perm), p(rows_el), p(rows_grad)))
    sh = (J.value + 1, n_chains)
    return dict(J=J.value, n_grad=ng.value, code1=code1.reshape(sh), code2=code2.reshape(sh), step1=step1.reshape(sh),
                step2=step2.reshape(sh), perm=perm, rows_el=rows_el, rows_grad=rows_grad)


def replay(pl, n_chains, nsw):
    """-> per chain list of events, in execution order."""
    ev = [[] for _ in range(n_chains)]

    def update(code, step, rows, j):
        for r in range(rows):
            op, c = int(code[j, r]), int(pl["perm"][r])
            if op & OP_LATCH:
                ev[c].append(("latch", int(step[j, r])))
            if op & OP_LATCH_CACHED:
                ev[c].append(("latch_cached", int(step[j, r])))
            if op & OP_POST:
                ev[c].append(("post", (op >> 8) & 15))
            if op & OP_PRE:
                ev[c].append(("pre", (op >> 12) & 15))

    for j in range(pl["J"] + 1):
        rows = int(pl["rows_el"][j])
        if j > 0:
            update(pl["code1"], pl["step1"], rows, j)
        for r in range(rows):
            if pl["code1"][j, r] & OP_FINISH:
                ev[int(pl["perm"][r])].append(("finish", int(pl["step1"][j, r])))
        if j == pl["J"]:
            break
        rg = int(pl["rows_grad"][j])
        began = False
        for r in range(rg):
            op = int(pl["code1"][j, r])
            if op & OP_BEGIN:
                began = True
                ev[int(pl["perm"][r])].append(("begin", int(pl["step1"][j, r]) + (1 if op & OP_FINISH else 0)))
        if began and j > 0:
            update(pl["code2"], pl["step2"], rg, j)
        for r in range(rg):
            ev[int(pl["perm"][r])].append(("grad", j))
        # rows beyond rg must have nothing left to do
        assert not pl["code1"][j + 1:, rg:].any() and not pl["code2"][j:, rg:].any()
    return ev


def expected(Lc, nsw):
    """The reference's event order for one chain with path lengths Lc (hmc.py:39-64), gradients at start points
    of transitions > 0 being reused."""
    out = []
    for t, L in enumerate(Lc):
        out.append(("begin", t))
        if t == 0:
            out.append(("grad", None))
            out.append(("latch", 0))
        else:
            out.append(("latch_cached", t))
        iters = max(L - 1, 0)
        first = True
        for it in range(iters):
            for v in range(nsw):
                if not first:
                    out.append(("post", (v - 1) % nsw))
                first = False
                out.append(("pre", v))
                out.append(("grad", None))
        if iters:
            out.append(("post", nsw - 1))
        else:
            out.extend([("grad", None)] * nsw)  # a transition without leapfrog iterations idles for one sweep
        out.append(("finish", t))
    return out


@pytest.mark.parametrize("nsw", [1, 2, 3])
@pytest.mark.parametrize("seed", [0, 1])
def test_stream_plan_event_order(nsw, seed):
    rs = np.random.RandomState(seed)
    n_steps, n_chains = 6, 9
    L = rs.randint(0, 7, size=(n_steps, n_chains)).astype(np.int32)  # includes L = 0 and L = 1 (no leapfrog iteration)
    L[2, 3] = 0
    L[0, 0] = 1
    pl = plan(L, nsw)
    ev = replay(pl, n_chains, nsw)
    launches_of = [[] for _ in range(n_chains)]
    for c in range(n_chains):
        got = [(k, (None if k == "grad" else v)) for k, v in ev[c]]
        assert got == expected(list(L[:, c]), nsw), "chain %d" % c
        launches_of[c] = [v for k, v in ev[c] if k == "grad"]
        assert launches_of[c] == list(range(len(launches_of[c])))  # a chain takes part in launches 0..T-1, no gaps
    # launch j > 0 follows a move of sweep group (j-1) % nsw for every row that moved at all
    for c in range(n_chains):
        pre_before = {}
        last_pre = None
        for k, v in ev[c]:
            if k == "pre":
                last_pre = v
            elif k == "grad":
                if last_pre is not None:
                    pre_before[v] = last_pre
                last_pre = None
        for j, v in pre_before.items():
            assert v == (j - 1) % nsw
    # bookkeeping: J = the longest chain, rows sorted by total work, evaluation count
    T = 1 + np.maximum(L - 1, 1).sum(axis=0) * nsw
    assert pl["J"] == T.max() and list(T[pl["perm"]]) == sorted(T, reverse=True)
    assert pl["n_grad"] == n_chains + (np.maximum(L - 1, 0).sum()) * nsw
    for j in range(pl["J"]):
        assert pl["rows_grad"][j] == (T > j).sum() and pl["rows_el"][j] == (T >= j).sum()


def test_stream_plan_single_step_and_uniform():
    pl = plan(np.full((1, 4), 3, np.int32), 2)  # one transition, L = 3: start point + 2 iterations x 2 groups
    assert pl["J"] == 5 and pl["n_grad"] == 4 * 5 and (pl["rows_grad"][:5] == 4).all()
    assert (pl["code1"][5] & OP_FINISH).all() and not (pl["code1"][5] & OP_BEGIN).any()
