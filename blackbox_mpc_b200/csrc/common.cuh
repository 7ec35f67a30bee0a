// common.cuh — context / model structures shared by the translation units of libbbmpc.so.
#pragma once
#include <cstdint>
#include <cstdio>
#include <string>
#include <vector>
#include <cuda_runtime.h>
#include "../../include/bbmpc.h"

namespace bbmpc {

constexpr int MAX_LAYERS = 8;
constexpr int MAX_MEMBERS = 16;
constexpr int MAX_DS = 64;
constexpr int MAX_DU = 32;
constexpr int BIAS_COLS = 3;  // bias enters the tensor-core GEMM as 3 bf16 rows times ones-columns

// Normalisation statistics as the kernels consume them (all device pointers into one buffer).
// den_* = std + 1e-7 rounded once in fp32, exactly the reference's (std + 1e-7) sub-expression
// (dynamics_handlers/system_dynamics_handler.py:119-124,153-156).
struct NormDev {
  int enabled;
  const float *mean_s, *den_s, *mean_a, *den_a, *mean_t, *den_t;
};

// One layer of the fp32 image used by the SIMT kernels: W32[K][ldw] (ldw = N rounded up to 16,
// zero padded), b32[ldw].  Offsets are in floats from MlpDev::w32, per member.
struct LayerDev {
  int K, N, ldw, act;
  int64_t w_off, b_off;  // member 0; member m adds m * MlpDev::w_member_stride
  // tensor-core image.  The N axis of a hidden layer is cut at column `nsplit` (a multiple of 32, 0 =
  // not cut) into two column ranges ("halves") that are contracted as separate MMA jobs, so that the
  // epilogue can convert the first range while the tensor pipe still works on the second.  Per
  // (member, layer) the image is [half 0: Kpad/16 chunks][half 1: Kpad/16 chunks]; a chunk of a half
  // with Nh columns holds [hi: 2 K-slabs x Nh x 16B][lo: 2 x Nh x 16B] (K-major 8x16B core matrices).
  int Kpad, Npad, nsplit;
  int64_t img_off;
};

// One Dense layer of one ensemble member as the MMA issuer of rollout_tc_kernel sees it; the jobs of
// one horizon step are listed in issue order (same order as the weight chunk table).
struct TcJob {
  uint32_t d_col, a_col;   // TMEM columns of the accumulator / of the A operand (hi of K-chunk 0)
  uint32_t idesc;          // tcgen05 instruction descriptor (M=128, N=Npad, bf16 x bf16 -> f32)
  uint32_t nchunks;        // K chunks of 16
  uint32_t desc_lo_base;   // low word of the smem matrix descriptor without the address field (LBO << 16)
  uint32_t lo_off16;       // (byte offset of the bf16-lo image inside a chunk) >> 4
  uint32_t flags;          // TCJ_*
  uint32_t chunk16;        // chunk_bytes >> 4: distance between consecutive K-chunks inside a ring stage
  uint32_t ngroups;        // the job's chunks reach shared memory in `ngroups` ring stages ...
  uint32_t gsz;            // ... of gsz[4*g +: 4] UNITS (pairs of K-chunks) each: one bulk copy, one full/empty barrier pair per group
  uint32_t pad[2];
};
constexpr int TC_GROUP_CAP_BYTES = 27648;  // ring-stage capacity target (one unit = 2 K-chunks of a 208-wide layer)
constexpr int TC_MAX_GROUPS = 8;
enum : uint32_t {
  TCJ_WAIT_X = 1u,        // first job of a step: wait until the epilogue has written the layer-0 input
  TCJ_FROM_EPI = 2u,      // A operand is produced chunk-wise by the epilogue (layer l >= 1)
  TCJ_ACC_FIRST = 4u,     // first MMA accumulates (output layer of ensemble member > 0)
  TCJ_ROUND_END = 8u,     // last job that reads the current A-operand round (flips the unit-barrier set)
  // completion barrier of the job: first-layer / later hidden layer x column half, or the output accumulator
  TCJ_COMMIT_D0H0 = 1u << 4, TCJ_COMMIT_D0H1 = 2u << 4, TCJ_COMMIT_DH0 = 3u << 4, TCJ_COMMIT_DH1 = 4u << 4,
  TCJ_COMMIT_DOUT = 5u << 4, TCJ_COMMIT_MASK = 7u << 4,
};

struct MlpDev {
  int n_members, n_layers;
  int64_t w_member_stride, img_member_stride;  // floats / bytes
  LayerDev layer[MAX_LAYERS];
  const float* w32;
  const uint8_t* wimg;   // nullptr when the model does not fit the tensor-core path
  const uint2* chunk_table;  // per step: (byte offset into wimg, bytes) of every chunk GROUP in issue order
  int chunks_per_step;       // number of groups per step
  int stage_bytes;           // largest group
  int early_l0;    // chunk-table order: first layer of member m+1 issued ahead of the output layer of member m
  const TcJob* jobs;   // [jobs_per_step]
  int jobs_per_step;
  // "solo" tables: the jobs / chunk groups of ONE member (member 0's image offsets; no early first-layer
  // issue, output layer not accumulated) — used when the members of an ensemble run on different CTAs.
  const uint2* solo_table; int solo_groups_per_step;
  const TcJob* solo_jobs; int solo_jobs_per_step;
  int max_width;   // widest activation (incl. input) — SIMT smem sizing
};

struct ModelHost {
  bool set = false;
  int dyn_id = BBMPC_DYN_MLP;
  int dS = 0, dU = 0;
  MlpDev mlp{};
  NormDev norm{};
  float* w32_buf = nullptr;
  uint8_t* wimg_buf = nullptr;
  uint2* chunk_table_buf = nullptr;
  TcJob* jobs_buf = nullptr;
  uint2* solo_table_buf = nullptr;
  TcJob* solo_jobs_buf = nullptr;
  float* norm_buf = nullptr;
  bool tc_ok = false;
  std::string tc_why;  // why the tensor-core path is unavailable for this model
};

}  // namespace bbmpc

struct bbmpc_ctx {
  int device = 0;
  uint64_t seed = 0;
  int sm_count = 148;
  int prec = BBMPC_PREC_AUTO;
  int reward_id = 0;
  bbmpc::ModelHost model;
  uint64_t launches = 0;
  uint64_t epoch = 1;        // bumped whenever model / statistics / reward / precision change: invalidates captured act() graphs
  std::string last_rollout_kernel;   // name of the kernel the last rollout launched (bench.py's roofline line)
  std::string err;
  // few-row step kernel (optimizer tail): per-member partial outputs + arrival counters
  float* step_scratch = nullptr;
  unsigned* step_counters = nullptr;
  // member-parallel rollout: per-group exchange of the members' raw outputs + arrival counters
  float* tc_xchg = nullptr; size_t tc_xchg_floats = 0;
  unsigned* tc_flags = nullptr; int tc_flags_n = 0;
  float* pipe_park = nullptr; size_t pipe_park_floats = 0;   // pipelined rollout: parked member-tile states
  bool tc_no_groups = false;   // a cooperative launch did not fit: stay with one CTA per tile
  // user reward compiled with NVRTC (user_reward.cu): loaded library, its two kernels, the source it was built from
  void* user_reward_lib = nullptr; cudaKernel_t user_reward_traj = nullptr; cudaKernel_t user_reward_rows = nullptr;
  int user_reward_dS = 0, user_reward_dU = 0; std::string user_reward_src;
  float* traj_buf = nullptr; size_t traj_floats = 0;   // visited states of the current rollout [rows][H][dS]
  float* traj_cur = nullptr;                             // non-null while a rollout must dump its states
  void* dbg_host = nullptr;  // BBMPC_DEBUG=1: host-mapped watchdog record of the tensor-core kernel
  // rollout-kernel timing (bbmpc_profile_*): event pairs recorded around every rollout launch
  bool prof_on = false;
  std::vector<cudaEvent_t> prof_ev;  // [2*i] start, [2*i+1] stop
  size_t prof_n = 0;                 // pairs recorded since the last read
};

namespace bbmpc {

int fail(bbmpc_ctx* ctx, int code, const char* fmt, ...);

#define BB_CUDA(ctx, expr)                                                                   \
  do {                                                                                       \
    cudaError_t e_ = (expr);                                                                 \
    if (e_ != cudaSuccess)                                                                   \
      return ::bbmpc::fail((ctx), BBMPC_ECUDA, "%s failed: %s (%s:%d)", #expr,               \
                           cudaGetErrorString(e_), __FILE__, __LINE__);                      \
  } while (0)

#define BB_LAUNCH_CHECK(ctx)                                                                 \
  do {                                                                                       \
    cudaError_t e_ = cudaGetLastError();                                                     \
    if (e_ != cudaSuccess)                                                                   \
      return ::bbmpc::fail((ctx), BBMPC_ECUDA, "kernel launch failed: %s (%s:%d)",           \
                           cudaGetErrorString(e_), __FILE__, __LINE__);                      \
    (ctx)->launches++;                                                                       \
  } while (0)

// ---- kernel launchers implemented per translation unit
struct StepIO {  // one predict/reward step over B rows (optimizer tail, predict_next_state, ...)
  const float* s; const float* a; const float* s2_in; float* s2_out; float* reward_out; float* raw_out;
  int B; int mode;  // mode bits: 1 = predict next state, 2 = reward, 4 = raw dynamics_function on x=s
};
int launch_rollout_simt(bbmpc_ctx* ctx, const float* states, const float* actions, float* returns,
                        const float* penalty, int rows, int A, int H, int act_ld, cudaStream_t st);
int launch_step_simt(bbmpc_ctx* ctx, const StepIO& io, cudaStream_t st);
int launch_rollout_tc(bbmpc_ctx* ctx, const float* states, const float* actions, float* returns,
                      const float* penalty, int rows, int A, int H, int passes, cudaStream_t st);
// pipelined two-jobs-in-flight variant (rollout_pipe.cu); returns -100 when the model does not fit it
int launch_rollout_pipe(bbmpc_ctx* ctx, const float* states, const float* actions, float* returns,
                        const float* penalty, int rows, int A, int H, int passes, cudaStream_t st);
int pack_tc_image(bbmpc_ctx* ctx, cudaStream_t st);   // builds model.wimg from model.w32
bool tc_supported(const ModelHost& m, std::string* why);
uint32_t tc_idesc(int Npad);  // instruction descriptor of the rollout kernel's MMAs
bool tc_column_map(const MlpDev& m, int* buf_w, int* col_x, int* col_dout);  // TMEM budget of the rollout kernel
int tc_du_slots(int dU);  // action slots at the head of the layer-0 K axis (8 or 16)
int resolve_precision(const bbmpc_ctx* ctx);
// user rewards (user_reward.cu)
void user_reward_free(bbmpc_ctx* ctx);
int user_reward_traj_buffer(bbmpc_ctx* ctx, int rows, int H, cudaStream_t st, float** out);
int launch_user_reward_traj(bbmpc_ctx* ctx, const float* traj, const float* states, const float* actions, const float* penalty,
                            float* returns, int rows, int A, int H, cudaStream_t st);
int launch_user_reward_rows(bbmpc_ctx* ctx, const float* s, const float* a, const float* s2, float* out, int B, cudaStream_t st);
// rollout dispatch used by bbmpc_rollout and the optimizers.  `penalty` (nullable, [rows]) is
// subtracted from the return before the NaN guard is applied?  No: the reference applies the NaN
// guard inside the evaluator and subtracts the penalty outside, so penalty is subtracted AFTER.
int rollout_dispatch(bbmpc_ctx* ctx, const float* states, const float* actions, float* returns,
                     const float* penalty, int rows, int A, int H, cudaStream_t st);

}  // namespace bbmpc
