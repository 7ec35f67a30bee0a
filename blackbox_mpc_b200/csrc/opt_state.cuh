// opt_state.cuh — the optimizer handle shared by optimizers.cu and cmaes.cu.
#pragma once
#include <vector>
#include "common.cuh"

struct bbmpc_opt {
  bbmpc_ctx* ctx = nullptr;
  bbmpc_opt_config cfg{};
  int rank = 0, world = 1;
  int p0 = 0, P_local = 0;        // this rank's slice [p0, p0 + P_local) of the population
  int n_eval = 0;                 // rows of the population axis handed to the evaluator (SPSA: 2*P_local)
  int HU = 0, AHU = 0;            // H*dU, A*H*dU
  bool cem_fuse = false, cem_fused_done = false;   // bbmpc_opt_call: CEM top-E + refit as one kernel (iter_merge then has nothing to do)
  int cem_slices = 1;             // population slices of the last unsharded CEM top-E (their messages sit back to back in d_partial)
  uint32_t act_call = 0;          // host mirror of *d_act_ctr
  uint32_t* d_act_ctr = nullptr;  // act() calls completed so far (Philox counter word), bumped by the last kernel of every call
  int time_step = 0;
  bool began = false;
  // device state
  float *d_lb = nullptr, *d_ub = nullptr;
  float* d_state = nullptr;       // [A,dS]
  float* d_samples = nullptr;     // [n_eval, A, H, dU]
  float* d_returns = nullptr;     // [n_eval, A]
  float* d_penalty = nullptr;     // [n_eval, A]
  float* d_mean = nullptr;        // loop variable  [A,H,dU]
  float* d_var = nullptr;         // loop variable  [A,H,dU] (CEM)
  float* d_prev = nullptr;        // persistent "previous solution" / current parameters [A,H,dU]
  float* d_var0 = nullptr;        // persistent solution variance [A,H,dU]
  float* d_partial = nullptr;     // [partial_floats]
  float* d_pi2_scratch = nullptr; // PI2: per-slice weighted sums [A, 64, H*dU]
  float* d_action = nullptr;      // [A,dU]
  float* d_next = nullptr;        // [A,dS]
  float* d_reward = nullptr;      // [A]
  // PSO
  float *d_v = nullptr, *d_pbx = nullptr, *d_pbr = nullptr, *d_gbx = nullptr, *d_gbr = nullptr, *d_sol = nullptr, *d_pso_r = nullptr;
  // CMA-ES
  float *d_m = nullptr, *d_sigma = nullptr, *d_C = nullptr, *d_B = nullptr, *d_D = nullptr, *d_ps = nullptr,
        *d_pc = nullptr, *d_z = nullptr, *d_BD = nullptr, *d_work = nullptr, *d_cma_w = nullptr;
  double cma_consts[16] = {0};
  // CMA-ES eigensolver state of THIS handle (cuSOLVER handle on the context's device, workspace, devInfo)
  void* eig_handle = nullptr; float* eig_work = nullptr; int eig_lwork = 0; int* eig_info = nullptr;
  // peer-memory exchange: [2 parities][partial_floats] messages + one sequence flag (last 64 bytes)
  float* p2p_buf = nullptr; size_t p2p_bytes = 0;
  float** d_peer = nullptr;        // [world] device table of the ranks' exchange buffers (own entry = p2p_buf)
  float* d_gather = nullptr;       // [world, partial_floats] local copy of the gathered messages
  std::vector<void*> p2p_opened;   // cudaIpcOpenMemHandle mappings to close
  bool p2p_on = false;
  uint32_t p2p_seq = 0;            // exchanges published so far (monotonic over the handle's lifetime)
  // trace + pinned staging
  float* trace = nullptr; int64_t trace_floats = 0;
  // injected standard variates (tests: committed golden draws): block k serves the k-th iteration since it was set
  const float* inject = nullptr; int64_t inject_floats = 0; int64_t inject_iter = 0;
  float* h_pinned = nullptr;
  // captured act(): the kernels of bbmpc_opt_call on the handle's own buffers, replayed with cudaGraphLaunch
  cudaGraphExec_t graph_exec = nullptr; uint64_t graph_epoch = 0; int graph_noise = -1; int graph_warm = 0;
  cudaStream_t graph_stream = nullptr;   // capture stream (the caller's stream may be the legacy default stream)
  uint64_t graph_launches = 0;   // kernels inside the captured graph (bbmpc_launch_count stays a kernel count)
  std::vector<void*> owned;
};


namespace bbmpc {
// CMA-ES (csrc/cmaes.cu): optimizers/cma_es.py behind the same begin / iter_local / iter_merge / finish protocol
int cmaes_create(bbmpc_opt* o);
void cmaes_destroy(bbmpc_opt* o);
int cmaes_set_shard(bbmpc_opt* o);
int cmaes_reset(bbmpc_opt* o, cudaStream_t st);
int cmaes_iter_local(bbmpc_opt* o, int iter, float* partial, cudaStream_t st);
int cmaes_iter_merge(bbmpc_opt* o, int iter, const float* partials, int world, cudaStream_t st);
// shared kernels implemented in optimizers.cu
void launch_penalty(const float* excess_sq, float* penalty, int64_t rows, int HU, cudaStream_t st);
void launch_topk_partial(const float* returns, const float* samples, float* partial, int P_local, int p0, int A, int HU, int E, cudaStream_t st,
                         int n_slices = 1, int64_t slice_stride = 0);
}  // namespace bbmpc
