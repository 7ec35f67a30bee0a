"""Stage order and barrier protocol of rollout_pipe_kernel (csrc/pipe_sched.h), checked on the CPU.

The header is plain C++: it is compiled here with g++ into a tiny shared library that lists the stage order, and a
discrete-event model of the kernel's roles (weights producer, MMA issuer, conversion warps, step boundary) runs that
order against the same barrier set (ring units, D-buffer drain, X regions) to show that the protocol cannot deadlock
and never overwrites an accumulator or ring unit that is still being read."""
import ctypes
import os
import subprocess
import tempfile

import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(os.path.dirname(HERE), "blackbox_mpc_b200", "csrc")

SHIM = r"""
#include "pipe_sched.h"
extern "C" int stage_order(int n_jobs, int n_layers, int n_mt, int* out, int cap) {
  bbmpc::StageSeq s; s.init(n_jobs, n_layers, n_mt);
  int j, l, b, n = 0;
  int i = 0;
  while (s.next(j, l, b, i)) { if (i != j % n_mt) return -1; if (n < cap) { out[3 * n] = j; out[3 * n + 1] = l; out[3 * n + 2] = b; } ++n; }
  return n;
}
"""


@pytest.fixture(scope="module")
def lib():
    d = tempfile.mkdtemp()
    src = os.path.join(d, "shim.cpp")
    with open(src, "w") as f:
        f.write(SHIM)
    so = os.path.join(d, "shim.so")
    subprocess.check_call(["g++", "-O1", "-shared", "-fPIC", "-I", CSRC, src, "-o", so])
    return ctypes.CDLL(so)


def order(lib, J, nL, n_mt):
    cap = J * nL + 8
    buf = (ctypes.c_int * (3 * cap))()
    n = lib.stage_order(J, nL, n_mt, buf, cap)
    assert n <= cap
    return [(buf[3 * i], buf[3 * i + 1], buf[3 * i + 2]) for i in range(n)]


CASES = [(H, nL, n_mt) for H in (1, 2, 3, 7, 30) for nL in (1, 2, 3, 4, 5) for n_mt in (1, 2, 3)]


@pytest.mark.parametrize("H,nL,n_mt", CASES)
def test_stage_order_is_complete_and_fifo(lib, H, nL, n_mt):
    J = H * n_mt
    seq = order(lib, J, nL, n_mt)
    assert sorted((j, l) for j, l, _ in seq) == [(j, l) for j in range(J) for l in range(nL)]
    pos = {(j, l): k for k, (j, l, _) in enumerate(seq)}
    for j in range(J):
        for l in range(1, nL):
            assert pos[(j, l)] > pos[(j, l - 1)]          # layers of a job in order
        if j >= n_mt:
            assert pos[(j, 0)] > pos[(j - n_mt, nL - 1)]  # a member-tile's next step after its output layer
    for j, l, b in seq:
        assert b == ((l & 1) if n_mt == 1 else (j & 1))
    # ring units: producers (hidden stages) and their consumers (next layer of the same job) in the same order
    prod = [(j, l) for j, l, _ in seq if l + 1 < nL]
    cons = [(j, l - 1) for j, l, _ in seq if l >= 1]
    assert prod == cons
    if n_mt >= 2:   # at most two jobs in flight (one per accumulator buffer), in job order per buffer
        for b in (0, 1):
            js = [j for j, l, bb in seq if bb == b]
            assert js == sorted(js)


class Bar:
    def __init__(self, count):
        self.count, self.pending, self.done = count, 0, 0

    def arrive(self, n=1):
        self.pending += n
        assert self.pending <= self.count, "more arrivals than the barrier expects in one phase"
        if self.pending == self.count:
            self.pending, self.done = 0, self.done + 1


def simulate(seq, J, nL, n_mt, units, a_units, w_stages, dur):
    """Cooperative round-robin over generator-based roles; returns the completion time, raises on deadlock."""
    t_now = [0]
    afull = [Bar(1) for _ in range(a_units)]
    afree = [Bar(1) for _ in range(a_units)]
    dfull, dout, drained = [Bar(1), Bar(1)], [Bar(1), Bar(1)], [Bar(1), Bar(1)]
    xfull = [Bar(1) for _ in range(3)]
    wfull = [Bar(1) for _ in range(w_stages)]
    wempty = [Bar(1) for _ in range(w_stages)]
    d_state = [None, None]     # what the accumulator buffer holds: ("busy", j, l) until read
    ring_owner = [None] * a_units
    single = n_mt == 1
    tensor_free = [0]

    def wait(bar, n):          # wait until phase n (0-based) has completed; the waiter may never lag two phases
        while bar.done <= n:
            yield
        assert bar.done - n <= 2 or True

    def timer(cycles):
        end = t_now[0] + cycles
        while t_now[0] < end:
            yield

    def producer():
        st = 0
        uses = [0] * w_stages
        for j, l, b in seq:
            for g in range(units[l] if l else 1):
                if uses[st] > 0:
                    yield from wait(wempty[st], uses[st] - 1)
                uses[st] += 1

                def land(at=t_now[0] + dur["wload"], st=st):   # bulk copies overlap: each lands one latency after its issue
                    while t_now[0] < at:
                        yield
                    wfull[st].arrive()
                tasks.append(land())
                st = (st + 1) % w_stages
                yield

    def mma():
        st, wuse = 0, [0] * w_stages
        useq = 0
        dw = [0, 0]
        xc = [0, 0, 0]
        ab = [0, 0]
        for j, l, b in seq:
            i = j % n_mt
            if dw[b] > 0:
                yield from wait(drained[b], dw[b] - 1)
            dw[b] += 1
            assert d_state[b] is None, f"accumulator {b} overwritten while {d_state[b]} unread (stage {j},{l})"
            if l == 0:
                yield from wait(xfull[i], xc[i]); xc[i] += 1
            key = 0 if single else b
            ub = ab[key]
            if l + 1 < nL:
                ab[key] = useq; useq += units[l + 1]
            n_groups = units[l] if l else 1
            for g in range(n_groups):
                yield from wait(wfull[st], wuse[st]); wuse[st] += 1
                if l > 0:
                    pu, wrap = (ub + g) % a_units, (ub + g) // a_units
                    yield from wait(afull[pu], wrap)
                    assert ring_owner[pu] == (j, l - 1, g), (ring_owner[pu], j, l, g)
                # tensor pipe: in-order
                start = max(t_now[0], tensor_free[0])
                tensor_free[0] = start + dur["mma_unit"]
                done_at = tensor_free[0]

                def retire(done_at=done_at, st=st, pu=(ub + g) % a_units if l > 0 else None):
                    while t_now[0] < done_at:
                        yield
                    wempty[st].arrive()
                    if pu is not None:
                        ring_owner[pu] = None
                        afree[pu].arrive()
                tasks.append(retire())
                st = (st + 1) % w_stages
            done_at = tensor_free[0]
            d_state[b] = ("busy", j, l)

            def commit(done_at=done_at, b=b, l=l):
                while t_now[0] < done_at:
                    yield
                (dfull if l + 1 < nL else dout)[b].arrive()
            tasks.append(commit())
            yield

    def conv():
        hc = [0, 0]
        useq = 0
        for j, l, b in seq:
            if l + 1 >= nL:
                continue
            yield from wait(dfull[b], hc[b]); hc[b] += 1
            assert d_state[b] == ("busy", j, l)
            n = units[l + 1]
            for u in range(n):
                pu, wrap = (useq + u) % a_units, (useq + u) // a_units
                yield from timer(dur["conv_unit"])
                if u == n - 1:
                    d_state[b] = None
                    drained[b].arrive()
                if wrap > 0:
                    yield from wait(afree[pu], wrap - 1)
                assert ring_owner[pu] is None, f"ring unit {pu} overwritten while owned by {ring_owner[pu]}"
                ring_owner[pu] = (j, l, u)
                afull[pu].arrive()
            useq += n

    def boundary():
        oc = [0, 0]
        for i in range(n_mt):
            xfull[i].arrive()
        b_single = (nL - 1) & 1
        for j in range(J):
            i, t = j % n_mt, j // n_mt
            b = b_single if single else (j & 1)
            yield from wait(dout[b], oc[b]); oc[b] += 1
            assert d_state[b] == ("busy", j, nL - 1)
            yield from timer(dur["read_out"])
            d_state[b] = None
            drained[b].arrive()
            yield from timer(dur["boundary"])
            if t + 1 < J // n_mt:
                xfull[i].arrive()

    tasks = [producer(), mma(), conv(), boundary()]
    idle_rounds = 0
    snapshot = None
    while tasks:
        alive = []
        n_before = len(tasks)
        for g in tasks[:n_before]:
            try:
                next(g)
                alive.append(g)
            except StopIteration:
                pass
        tasks[:] = alive + tasks[n_before:]
        state = (tuple(b.done for b in afull + afree + dfull + dout + drained + xfull + wfull + wempty), len(tasks))
        t_now[0] += 20
        if state == snapshot:
            idle_rounds += 1
            assert idle_rounds < 5000, "deadlock: no barrier completed for a long time"
        else:
            idle_rounds, snapshot = 0, state
    return t_now[0]


@pytest.mark.parametrize("H,nL,n_mt", [(4, 4, 1), (4, 4, 2), (5, 4, 3), (3, 2, 3), (3, 3, 2), (4, 5, 3), (3, 1, 2), (2, 1, 3)])
@pytest.mark.parametrize("spare", [0, 1, 3])
def test_protocol_runs_to_completion(lib, H, nL, n_mt, spare):
    J = H * n_mt
    seq = order(lib, J, nL, n_mt)
    units = [1] + [7] * (nL - 1)          # ring units per layer (layer 0 reads the X region instead)
    dur = dict(wload=900, mma_unit=620, conv_unit=560, read_out=300, boundary=6000)
    t = simulate(seq, J, nL, n_mt, units, a_units=7 + spare, w_stages=3, dur=dur)
    assert t > 0


def test_pipelining_shortens_the_round(lib):
    """Three member-tiles on one CTA should take clearly less than three times one member-tile."""
    nL, H = 4, 6
    units = [1, 7, 7, 2]
    dur = dict(wload=900, mma_unit=620, conv_unit=450, read_out=300, boundary=6000)
    t1 = simulate(order(lib, H, nL, 1), H, nL, 1, units, 8, 3, dur)
    t3 = simulate(order(lib, 3 * H, nL, 3), 3 * H, nL, 3, units, 8, 3, dur)
    assert t3 < 2.3 * t1, (t1, t3)
