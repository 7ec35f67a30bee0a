/* bbmpc.h — C ABI of the B200-native sampling-MPC rollout engine (libbbmpc.so).
 *
 * The reference (ossamaAhmed/blackbox_mpc @ 68c9e63) is pure Python/TensorFlow and has no FFI;
 * each entry point below replaces the TF graph behind one reference Python interface (cited as
 * blackbox_mpc/<file>:<line>) and is what blackbox_mpc_b200's Python classes bind via ctypes.
 *
 * Conventions
 *   - plain C types only; `stream` is a cudaStream_t passed as void* (NULL = default stream);
 *   - every `const float*` / `float*` is a DEVICE pointer, fp32, C-contiguous, unless the
 *     parameter name ends in `_host`;
 *   - calls enqueue work on `stream` and return without synchronising unless stated;
 *   - return 0 on success, a negative BBMPC_E* code on failure; bbmpc_last_error() gives text;
 *   - there is NO CPU fallback: without a CUDA device every compute call fails with BBMPC_ECUDA.
 *
 * Shapes: P population, A num_agents, H planning_horizon, dS/dU state/action dims.
 */
#ifndef BBMPC_H_
#define BBMPC_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define BBMPC_OK 0
#define BBMPC_EINVAL (-1)   /* bad argument / unsupported shape */
#define BBMPC_ECUDA (-2)    /* CUDA runtime error (text in bbmpc_last_error) */
#define BBMPC_ESTATE (-3)   /* call order violated (e.g. rollout before a model is set) */
#define BBMPC_ENOMEM (-4)

/* dynamics ids (bbmpc_model_set_builtin) */
#define BBMPC_DYN_MLP 0       /* set through bbmpc_model_set_mlp */
#define BBMPC_DYN_PENDULUM 1  /* utils/pendulum.py:58-92 PendulumTrueModel (returns deviation) */

/* reward ids (bbmpc_reward_set_builtin) */
#define BBMPC_REWARD_PENDULUM 1     /* utils/pendulum.py:10-35 called as (s, a, s') by deterministic.py:65-66
                                       -> its `actions` argument receives next_state (reference behaviour) */
#define BBMPC_REWARD_HALFCHEETAH 2  /* tutorials/mujoco/cost_func.py:5-22 */
#define BBMPC_REWARD_PENDULUM_GYM 3 /* pendulum reward with the arguments as its docstring intends (s, s', a) */
#define BBMPC_REWARD_USER 4         /* CUDA source supplied through bbmpc_reward_set_nvrtc */

/* activation ids for bbmpc_model_set_mlp */
#define BBMPC_ACT_NONE 0
#define BBMPC_ACT_TANH 1
#define BBMPC_ACT_RELU 2
#define BBMPC_ACT_SIGMOID 3

/* arithmetic of the MLP contraction (bbmpc_set_precision) */
#define BBMPC_PREC_AUTO 0    /* BF16X3 tensor-core path when the model fits it, else FP32 SIMT */
#define BBMPC_PREC_FP32 1    /* fp32 FFMA on CUDA cores (parity-grade, any layer width) */
#define BBMPC_PREC_BF16X3 2  /* tcgen05, operands split hi+lo bf16, 3 MMAs, fp32 accumulate */
#define BBMPC_PREC_BF16 3    /* tcgen05, single bf16 pass (NOT fp32-grade; opt-in only) */

/* optimizer kinds */
#define BBMPC_OPT_CEM 1
#define BBMPC_OPT_PI2 2
#define BBMPC_OPT_RANDOM_SEARCH 3
#define BBMPC_OPT_PSO 4
#define BBMPC_OPT_SPSA 5
#define BBMPC_OPT_CMAES 6

typedef struct bbmpc_ctx bbmpc_ctx;
typedef struct bbmpc_opt bbmpc_opt;

/* ---- library / context ------------------------------------------------------------------ */
int bbmpc_version(void);
/* sizeof(bbmpc_opt_config) as the library was compiled: bindings compare it with their own mirror of the struct
 * so that a stale libbbmpc.so fails loudly instead of reading a config with shifted fields. */
int bbmpc_abi_config_size(void);
/* Text of the last error on this ctx (or of the last failed bbmpc_ctx_create when ctx == NULL). */
const char* bbmpc_last_error(const bbmpc_ctx* ctx);
/* One context per process per GPU.  `seed` keys every Philox stream drawn by this context. */
int bbmpc_ctx_create(int device, uint64_t seed, bbmpc_ctx** out);
void bbmpc_ctx_destroy(bbmpc_ctx* ctx);
int bbmpc_set_precision(bbmpc_ctx* ctx, int prec);
/* Which BBMPC_PREC_* the next rollout will actually use (AUTO resolved). */
int bbmpc_get_effective_precision(const bbmpc_ctx* ctx);
/* Number of kernels this context has launched so far (bench.py's gpu_launches). */
uint64_t bbmpc_launch_count(const bbmpc_ctx* ctx);
/* Name of the kernel the last rollout launched (rollout_pipe_kernel / rollout_tc_kernel / rollout_simt_kernel). */
const char* bbmpc_last_rollout_kernel(const bbmpc_ctx* ctx);
/* Rollout-kernel timing for bench.py's roofline: while enabled, every rollout launch is bracketed
 * by a CUDA event pair on its own stream.  bbmpc_profile_read synchronises on the recorded events,
 * returns the summed device time (ms) and the number of launches, and clears the record. */
int bbmpc_profile_enable(bbmpc_ctx* ctx, int on);
int bbmpc_profile_read(bbmpc_ctx* ctx, double* ms_total, int64_t* n_launches);

/* ---- dynamics model: dynamics_functions/deterministic_mlp.py:20-24,49-51 (Dense chain) and
 *      dynamics_handlers/system_dynamics_handler.py:97-161 (process_input / process_output) ---- */
/* n_members MLPs (an ensemble averages the members' raw outputs, member order, then / n).
 * dims[n_layers+1]; W[m*n_layers + l] -> [dims[l], dims[l+1]] row-major (Keras Dense kernel
 * layout), b[...] -> [dims[l+1]]; act_ids[n_layers].  Weights are copied/re-packed into the
 * context (call again after re-training: the handler object is shared with the trainer,
 * utils/iterative_mpc.py:147-157). */
int bbmpc_model_set_mlp(bbmpc_ctx* ctx, int n_members, int n_layers, const int* dims,
                        const float* const* W, const float* const* b, const int* act_ids,
                        void* stream);
/* Normalisation statistics mean/std of states [dS], actions [dU], targets [dS]
 * (system_dynamics_handler.py:119-124,153-156).  All six NULL => is_normalized=False. */
int bbmpc_model_set_norm(bbmpc_ctx* ctx, int dS, int dU, const float* mean_s, const float* std_s,
                         const float* mean_a, const float* std_a, const float* mean_t,
                         const float* std_t, void* stream);
/* Analytical true model (true_model=True: raw concat in, s' = s + f(x) out). */
int bbmpc_model_set_builtin(bbmpc_ctx* ctx, int dyn_id, int dS, int dU);
int bbmpc_reward_set_builtin(bbmpc_ctx* ctx, int reward_id);
/* User-supplied reward_function (policies/mpc_policy.py:42-44; e.g. tutorials/mujoco/cost_func.py:5-22) as CUDA source
 * defining  __device__ float reward(const float* s, const float* a, const float* s2)  over dS / dU / dS floats
 * (BBMPC_DS and BBMPC_DU are predefined macros).  Compiled once per context with NVRTC (sm_100a, --fmad=false: plain
 * IEEE fp32 operations); every rollout then sums it over the visited states, every bbmpc_reward call applies it row-wise.
 * Needs a model (dS, dU) to be set first.  Compilation errors return BBMPC_EINVAL with the NVRTC log as error text. */
int bbmpc_reward_set_nvrtc(bbmpc_ctx* ctx, const char* cuda_source);

/* ---- evaluator: trajectory_evaluators/deterministic.py ----------------------------------- */
/* __call__ (:26-77): returns[P,A] = sum_t reward(s_t, a_t, s_{t+1}), NaN -> -1e6.
 * states [A,dS]; actions [P,A,H,dU] (row p*A+a starts from states[a]). */
int bbmpc_rollout(bbmpc_ctx* ctx, const float* states, const float* actions, float* returns, int P,
                  int A, int H, void* stream);
/* predict_next_state (:79-103): s[B,dS], a[B,dU] -> out[B,dS]. */
int bbmpc_predict_next_state(bbmpc_ctx* ctx, const float* s, const float* a, float* out, int B,
                             void* stream);
/* evaluate_next_reward (:105-127) = reward_function(s, a, s2): out[B]. */
int bbmpc_reward(bbmpc_ctx* ctx, const float* s, const float* a, const float* s2, float* out, int B,
                 void* stream);
/* dynamics_function(x, train=False) itself (deterministic_mlp.py:27-51 / pendulum.py:58-92):
 * x[B,dS+dU] -> out[B,dS], no handler pre/post-processing. */
int bbmpc_dynamics_forward(bbmpc_ctx* ctx, const float* x, float* out, int B, void* stream);

/* ---- optimizers: optimizers/optimizer_base.py:6-115 + one subclass each -------------------
 * Hyper-parameters mirror the reference constructors.  lb/ub are HOST arrays [dU]
 * (env_action_space.low/.high).  Optimizer state (tf.Variables in the reference) lives on the
 * device, owned by the handle. */
typedef struct bbmpc_opt_config {
  int kind;             /* BBMPC_OPT_* */
  int population_size;  /* P (global, before sharding) */
  int num_agents;       /* A */
  int planning_horizon; /* H */
  int max_iterations;   /* ignored by RandomSearch (single shot) */
  int dS, dU;
  const float* lb_host; /* [dU] */
  const float* ub_host; /* [dU] */
  int num_elite;        /* CEM cem.py:8-10, CMA-ES cma_es.py:8-10 */
  float alpha;          /* CEM smoothing (cem.py:10) | SPSA alpha (spsa.py:9) */
  float epsilon;        /* CEM: stored, never used (cem.py:53) */
  float lamda;          /* PI2 (pi2.py:11) */
  float c1, c2, w, initial_velocity_fraction; /* PSO (pso.py:8-11) */
  float gamma, a_par, noise_parameter;        /* SPSA (spsa.py:10-12) */
  float h_sigma, alpha_cov;                   /* CMA-ES (cma_es.py:9-10) */
} bbmpc_opt_config;

int bbmpc_opt_create(bbmpc_ctx* ctx, const bbmpc_opt_config* cfg, bbmpc_opt** out);
void bbmpc_opt_destroy(bbmpc_opt* opt);
/* reset(): per-subclass semantics of the reference (e.g. cem.py:138-149 resets the mean only). */
int bbmpc_opt_reset(bbmpc_opt* opt, void* stream);
/* Population sharding for one-process-per-GPU runs: this rank evaluates global rows
 * [rank*P/world, (rank+1)*P/world).  Samples are keyed on the GLOBAL row, so results do not
 * depend on `world`.  Default rank 0 of 1. */
int bbmpc_opt_set_shard(bbmpc_opt* opt, int rank, int world);

/* OptimizerBase.__call__ (optimizer_base.py:55-95) on one GPU: _optimize, optional exploration
 * noise + clip, predict_next_state, evaluate_next_reward.  state[A,dS] -> action[A,dU],
 * next_state[A,dS], reward[A].  Fully asynchronous on `stream`. */
int bbmpc_opt_call(bbmpc_opt* opt, const float* state, int time_step, int add_exploration_noise,
                   float* action, float* next_state, float* reward, void* stream);
/* Same with HOST buffers (MPCPolicy.act, policies/mpc_policy.py:149-172): H2D of the observation,
 * the whole call, D2H of the three results through pinned staging, then synchronises. */
int bbmpc_opt_call_host(bbmpc_opt* opt, const float* state_host, int time_step,
                        int add_exploration_noise, float* action_host, float* next_state_host,
                        float* reward_host, void* stream);

/* Split form for sharded populations (one small exchange per iteration):
 *   begin; for it in range(n_iters): iter_local -> [all_gather partials] -> iter_merge; finish.
 * partial_floats() is the per-rank message size; iter_merge takes `world` messages back to back. */
int bbmpc_opt_num_iterations(const bbmpc_opt* opt);
int bbmpc_opt_partial_floats(const bbmpc_opt* opt);
int bbmpc_opt_begin(bbmpc_opt* opt, const float* state, int time_step, void* stream);
int bbmpc_opt_iter_local(bbmpc_opt* opt, int iter, float* partial_out, void* stream);
int bbmpc_opt_iter_merge(bbmpc_opt* opt, int iter, const float* partials, int world, void* stream);
int bbmpc_opt_finish(bbmpc_opt* opt, int add_exploration_noise, float* action, float* next_state,
                     float* reward, void* stream);

/* Peer-memory exchange (one process per GPU on one NVLink/NVSwitch box; new: the reference is single-device).
 * Instead of an NCCL all_gather between iter_local and iter_merge, every rank publishes its partial message
 * in a device buffer that the other ranks have mapped (CUDA IPC), and the merge side of each rank pulls the
 * peers' messages over NVLink inside one kernel (flag wait + P2P loads).  Once connected, bbmpc_opt_call /
 * bbmpc_opt_call_host run the whole sharded act() without returning to the host between iterations.
 *   bbmpc_opt_p2p_export: allocates the exchange buffer and writes its 64-byte cudaIpcMemHandle_t to
 *                         handle_out (host) and its device address to ptr_out (for same-process peers).
 *   bbmpc_opt_p2p_connect: handles_host = world x 64 bytes (rank order) or NULL; ptrs_host = world device
 *                         addresses (same-process peers / tests) or NULL.  Entry `rank` is ignored. */
int bbmpc_opt_p2p_export(bbmpc_opt* opt, void* handle_out_host, void** ptr_out_host);
int bbmpc_opt_p2p_connect(bbmpc_opt* opt, const void* handles_host, void* const* ptrs_host);

/* Test/inspection hooks.  name: "mean" "variance" (CEM/PI2), "solution" (SPSA), "samples"
 * (last iteration's local samples [P_local,A,H,dU]), "returns" ([P_local,A], penalties applied),
 * "m" "sigma" "C" "B" "D" (diagonal) "p_sigma" "p_C" (CMA-ES), "x" "v" "pbest_x" "pbest_r" "gbest_x" (PSO).
 * Copies min(n_floats, size) floats into out (device) and returns the tensor's size in floats. */
int64_t bbmpc_opt_get_tensor(bbmpc_opt* opt, const char* name, float* out, int64_t n_floats,
                             void* stream);
/* Test hook (tests/test_gpu_golden.py): replace the optimizer's in-kernel Philox draws by injected STANDARD variates
 * (truncated-normal z for CEM / PI2, U[0,1) for RandomSearch, +-1 for SPSA, N(0,1) for CMA-ES), fp32 on the device,
 * one block of [population_size, num_agents, H*dU] floats (CMA-ES: [population_size, A*H*dU]) per optimizer
 * iteration since this call, indexed by GLOBAL population row.  NULL switches back to Philox.  Not available for PSO. */
int bbmpc_opt_set_draw_injection(bbmpc_opt* opt, const float* std_draws, int64_t n_floats);
/* When set (device buffer of n_iters*P_local*A*H*dU floats), every iteration's evaluated samples
 * are also recorded there, so a test can inject the exact draws into the oracle.  NULL disables.
 * CMA-ES records its raw N(0,1) draws z [P_local, N] per iteration instead (the samples follow from z
 * through the solver-specific eigenbasis B). */
int bbmpc_opt_set_sample_trace(bbmpc_opt* opt, float* trace, int64_t n_floats);

/* ---- sampler known-answer hooks (host-side, no GPU needed) -------------------------------- */
/* Philox4x32-10 block function used by every in-kernel sampler. */
void bbmpc_philox4x32_host(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4]);

#ifdef __cplusplus
}
#endif
#endif /* BBMPC_H_ */
