// pipe_sched.h — static stage order of rollout_pipe_kernel (host + device; plain C++ so that the CPU
// test-suite can compile it with g++ and simulate the barrier protocol, tests/test_pipe_schedule_cpu.py).
//
// A JOB is one horizon step of one member-tile; job j of a CTA with n_mt member-tiles is member-tile
// j % n_mt at step j / n_mt.  A STAGE is one Dense layer of a job (l = 0 .. nL-1).
//
//   n_mt == 1 ("single"):   stages in plain order; the accumulator of layer l lives in D[l & 1], so the
//                           conversion of layer l overlaps the contraction of layer l+1 chunk by chunk.
//   n_mt >= 2 ("pipelined"): job j owns D[j & 1] for all its layers.  The stages of the two slots alternate
//                           group by group; slot 0 leads by `lead` groups.  With n_mt >= 3 the output layer
//                           of job j and layer 0 of job j+2 (the next job on that slot, a different
//                           member-tile whose input is ready) form one group; with n_mt == 2 the next job on
//                           the slot is the SAME member-tile and needs the step boundary of job j first, so
//                           its layer 0 is a group of its own, half a job later.
// In every mode the hidden stages (producers of A-operand ring units) and their consumers (the next layer of
// the same job) appear in the same relative order: ring units are produced and consumed FIFO.
#pragma once

#if defined(__CUDACC__)
#define PIPE_HD __host__ __device__
#else
#define PIPE_HD
#endif

namespace bbmpc {

struct StageSeq {
  int nL, single, merged, nmt;
  // per slot: next stage index / number of stages (stage e: job = slot + 2 * (e / nL), layer = e % nL); scalars, not
  // arrays, so that the device code keeps them in registers.  q / r / t: e / nL, e % nL and (job % n_mt) of the slot's
  // current job, carried along instead of divided out (the walkers are latency-critical single warps).
  int e0, e1, n0, n1;
  int q0, q1, r0, r1, t0, t1;
  int turn, lead, grp_left, cur;

  PIPE_HD void init(int n_jobs, int n_layers, int n_mt) {
    nL = n_layers;
    nmt = n_mt;
    single = (n_mt == 1);
    merged = (n_mt >= 3 && n_layers >= 2);
    e0 = e1 = 0;
    q0 = q1 = r0 = r1 = 0;
    t0 = 0; t1 = (n_mt > 1) ? 1 : 0;
    if (single) { n0 = nL * n_jobs; n1 = 0; }
    else { n0 = nL * ((n_jobs + 1) / 2); n1 = nL * (n_jobs / 2); }
    // slot 0 runs `lead` groups before slot 1 starts (about half a job: the jobs then start and finish in job order
    // and the step boundary of job j is over when layer 0 of job j + n_mt is due), then the slots alternate, slot 1 first
    turn = 1;
    lead = (merged ? (nL - 1) / 2 : nL / 2) + 1;
    grp_left = 0;
    cur = 0;
  }

  // Next stage in issue order: job j, layer l, accumulator buffer b, member-tile slot i = j % n_mt of the CTA.
  // false when the round is complete.
  PIPE_HD bool next(int& j, int& l, int& b, int& i) {
    if (single) {
      if (e0 >= n0) return false;
      j = q0; l = r0; b = l & 1; i = 0;
      ++e0;
      if (++r0 == nL) { r0 = 0; ++q0; }
      return true;
    }
    if (grp_left == 0) {
      int s;
      if (lead > 0) { s = 0; --lead; } else { s = turn; turn ^= 1; }
      if ((s ? e1 : e0) >= (s ? n1 : n0)) s ^= 1;
      const int es = s ? e1 : e0, ns = s ? n1 : n0;
      if (es >= ns) return false;
      cur = s;
      grp_left = (merged && (s ? r1 : r0) == nL - 1 && es + 1 < ns) ? 2 : 1;
    }
    const int s = cur;
    b = s;
    if (s) {
      j = 1 + 2 * q1; l = r1; i = t1;
      ++e1;
      if (++r1 == nL) { r1 = 0; ++q1; t1 += 2; if (t1 >= nmt) t1 -= nmt; }
    } else {
      j = 2 * q0; l = r0; i = t0;
      ++e0;
      if (++r0 == nL) { r0 = 0; ++q0; t0 += 2; if (t0 >= nmt) t0 -= nmt; }
    }
    --grp_left;
    return true;
  }
  PIPE_HD bool next(int& j, int& l, int& b) { int i; return next(j, l, b, i); }
};

}  // namespace bbmpc
