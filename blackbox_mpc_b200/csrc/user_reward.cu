// user_reward.cu — reward functions supplied as CUDA source (bbmpc_reward_set_nvrtc).
//
// The reference's plug point is an arbitrary callable reward_function(current_state, actions, next_state)
// (policies/mpc_policy.py:42-44, invoked at trajectory_evaluators/deterministic.py:65-66 and :126-127); the
// headline tutorial's reward is user code (tutorials/mujoco/cost_func.py:5-22).  A Python callable cannot run
// inside the rollout kernel, a CUDA one can: the user supplies
//
//     __device__ float reward(const float* s, const float* a, const float* s2)      // dS, dU, dS floats
//
// (BBMPC_DS / BBMPC_DU are defined as macros).  It is compiled once with NVRTC into two small kernels:
//   * bbmpc_user_reward_traj: per trajectory, sum of reward(s_t, a_t, s_{t+1}) over the horizon in step order,
//     NaN -> -1e6 (deterministic.py:75-77), minus the optimizer's penalty.  The rollout kernels run with their
//     own reward switched off and DUMP the visited states ([row][t][dS], one extra HBM pass of rows*H*dS*4 B:
//     24 MB for the C4 configuration, microseconds) — so a user reward runs on the tensor-core path at full speed
//     and is not welded into the 900-line kernels;
//   * bbmpc_user_reward_rows: reward of B (s, a, s2) rows (evaluate_next_reward, optimizer_base.py:93-94).
// Compiled with --fmad=false: every + - * / is the IEEE fp32 operation the reference's TF graph performs.
// NVRTC is bound lazily with dlopen (libbbmpc.so does not link it), like cuSOLVER in cmaes.cu.
#include <dlfcn.h>
#include <mutex>
#include <string>
#include <vector>
#include "common.cuh"

namespace bbmpc {
namespace {

struct Nvrtc {
  void* lib = nullptr;
  int (*create)(void**, const char*, const char*, int, const char* const*, const char* const*) = nullptr;
  int (*compile)(void*, int, const char* const*) = nullptr;
  int (*log_size)(void*, size_t*) = nullptr;
  int (*get_log)(void*, char*) = nullptr;
  int (*cubin_size)(void*, size_t*) = nullptr;
  int (*get_cubin)(void*, char*) = nullptr;
  int (*destroy)(void**) = nullptr;
  bool ok = false;
};
Nvrtc g_nvrtc;
std::once_flag g_nvrtc_once;

void nvrtc_bind() {
  Nvrtc& n = g_nvrtc;
  const char* names[] = {"libnvrtc.so.12", "libnvrtc.so", "/usr/local/cuda/lib64/libnvrtc.so.12", "/usr/local/cuda/lib64/libnvrtc.so"};
  for (const char* s : names) { n.lib = dlopen(s, RTLD_NOW | RTLD_LOCAL); if (n.lib) break; }
  if (!n.lib) return;
  n.create = reinterpret_cast<decltype(n.create)>(dlsym(n.lib, "nvrtcCreateProgram"));
  n.compile = reinterpret_cast<decltype(n.compile)>(dlsym(n.lib, "nvrtcCompileProgram"));
  n.log_size = reinterpret_cast<decltype(n.log_size)>(dlsym(n.lib, "nvrtcGetProgramLogSize"));
  n.get_log = reinterpret_cast<decltype(n.get_log)>(dlsym(n.lib, "nvrtcGetProgramLog"));
  n.cubin_size = reinterpret_cast<decltype(n.cubin_size)>(dlsym(n.lib, "nvrtcGetCUBINSize"));
  n.get_cubin = reinterpret_cast<decltype(n.get_cubin)>(dlsym(n.lib, "nvrtcGetCUBIN"));
  n.destroy = reinterpret_cast<decltype(n.destroy)>(dlsym(n.lib, "nvrtcDestroyProgram"));
  n.ok = n.create && n.compile && n.log_size && n.get_log && n.cubin_size && n.get_cubin && n.destroy;
}

const char* kWrapper = R"SRC(
extern "C" __global__ void bbmpc_user_reward_traj(const float* __restrict__ traj, const float* __restrict__ states0,
                                                  const float* __restrict__ actions, const float* __restrict__ penalty,
                                                  float* __restrict__ returns, int rows, int A, int H) {
  const int row = blockIdx.x * blockDim.x + threadIdx.x;
  if (row >= rows) return;
  float s[BBMPC_DS], s2[BBMPC_DS], a[BBMPC_DU];
  for (int k = 0; k < BBMPC_DS; ++k) s[k] = states0[(row % A) * BBMPC_DS + k];
  float ret = 0.0f;
  for (int t = 0; t < H; ++t) {
    const float* tp = traj + ((size_t)row * H + t) * BBMPC_DS;
    const float* ap = actions + ((size_t)row * H + t) * BBMPC_DU;
    for (int k = 0; k < BBMPC_DS; ++k) s2[k] = tp[k];
    for (int k = 0; k < BBMPC_DU; ++k) a[k] = ap[k];
    ret = ret + reward(s, a, s2);
    for (int k = 0; k < BBMPC_DS; ++k) s[k] = s2[k];
  }
  float r = (ret != ret) ? -1e6f : ret;
  if (penalty) r = r - penalty[row];
  returns[row] = r;
}
extern "C" __global__ void bbmpc_user_reward_rows(const float* __restrict__ s_in, const float* __restrict__ a_in,
                                                  const float* __restrict__ s2_in, float* __restrict__ out, int B) {
  const int row = blockIdx.x * blockDim.x + threadIdx.x;
  if (row >= B) return;
  float s[BBMPC_DS], s2[BBMPC_DS], a[BBMPC_DU];
  for (int k = 0; k < BBMPC_DS; ++k) { s[k] = s_in[(size_t)row * BBMPC_DS + k]; s2[k] = s2_in[(size_t)row * BBMPC_DS + k]; }
  for (int k = 0; k < BBMPC_DU; ++k) a[k] = a_in[(size_t)row * BBMPC_DU + k];
  out[row] = reward(s, a, s2);
}
)SRC";

}  // namespace

void user_reward_free(bbmpc_ctx* ctx) {
  if (ctx->user_reward_lib) cudaLibraryUnload(static_cast<cudaLibrary_t>(ctx->user_reward_lib));
  ctx->user_reward_lib = nullptr; ctx->user_reward_traj = nullptr; ctx->user_reward_rows = nullptr;
  cudaFree(ctx->traj_buf); ctx->traj_buf = nullptr; ctx->traj_floats = 0;
}

int user_reward_compile(bbmpc_ctx* ctx, const char* src) {
  std::call_once(g_nvrtc_once, nvrtc_bind);
  Nvrtc& n = g_nvrtc;
  if (!n.ok) return fail(ctx, BBMPC_ECUDA, "bbmpc_reward_set_nvrtc needs libnvrtc.so.12 (dlopen failed)");
  const ModelHost& m = ctx->model;
  if (!m.dS || !m.dU) return fail(ctx, BBMPC_ESTATE, "set a dynamics model (dS, dU) before a user reward");
  std::string code = "#define BBMPC_DS " + std::to_string(m.dS) + "\n#define BBMPC_DU " + std::to_string(m.dU) + "\n";
  code += src;
  code += kWrapper;
  void* prog = nullptr;
  if (n.create(&prog, code.c_str(), "bbmpc_user_reward.cu", 0, nullptr, nullptr) != 0)
    return fail(ctx, BBMPC_ECUDA, "nvrtcCreateProgram failed");
  const char* opts[] = {"--gpu-architecture=sm_100a", "--fmad=false", "--std=c++17", "-default-device"};
  const int rc = n.compile(prog, 4, opts);
  if (rc != 0) {
    size_t ls = 0; n.log_size(prog, &ls);
    std::string log(ls ? ls : 1, '\0');
    if (ls) n.get_log(prog, &log[0]);
    n.destroy(&prog);
    return fail(ctx, BBMPC_EINVAL, "user reward does not compile:\n%s", log.c_str());
  }
  size_t cs = 0; n.cubin_size(prog, &cs);
  std::vector<char> cubin(cs);
  n.get_cubin(prog, cubin.data());
  n.destroy(&prog);
  user_reward_free(ctx);
  cudaLibrary_t lib = nullptr;
  BB_CUDA(ctx, cudaLibraryLoadData(&lib, cubin.data(), nullptr, nullptr, 0, nullptr, nullptr, 0));
  cudaKernel_t k1 = nullptr, k2 = nullptr;
  BB_CUDA(ctx, cudaLibraryGetKernel(&k1, lib, "bbmpc_user_reward_traj"));
  BB_CUDA(ctx, cudaLibraryGetKernel(&k2, lib, "bbmpc_user_reward_rows"));
  ctx->user_reward_lib = lib; ctx->user_reward_traj = k1; ctx->user_reward_rows = k2;
  ctx->user_reward_dS = m.dS; ctx->user_reward_dU = m.dU;
  return BBMPC_OK;
}

// Trajectory buffer of the current rollout: [rows][H][dS] floats.
int user_reward_traj_buffer(bbmpc_ctx* ctx, int rows, int H, cudaStream_t st, float** out) {
  const size_t need = static_cast<size_t>(rows) * H * ctx->model.dS;
  if (ctx->traj_floats < need) {
    BB_CUDA(ctx, cudaStreamSynchronize(st));
    cudaFree(ctx->traj_buf);
    ctx->traj_buf = nullptr; ctx->traj_floats = 0;
    BB_CUDA(ctx, cudaMalloc(&ctx->traj_buf, need * sizeof(float)));
    ctx->traj_floats = need;
  }
  *out = ctx->traj_buf;
  return BBMPC_OK;
}

int launch_user_reward_traj(bbmpc_ctx* ctx, const float* traj, const float* states, const float* actions, const float* penalty,
                            float* returns, int rows, int A, int H, cudaStream_t st) {
  if (!ctx->user_reward_traj) return fail(ctx, BBMPC_ESTATE, "no user reward is compiled");
  if (ctx->user_reward_dS != ctx->model.dS || ctx->user_reward_dU != ctx->model.dU)
    return fail(ctx, BBMPC_ESTATE, "the user reward was compiled for dS=%d dU=%d", ctx->user_reward_dS, ctx->user_reward_dU);
  void* args[] = {&traj, &states, &actions, &penalty, &returns, &rows, &A, &H};
  BB_CUDA(ctx, cudaLaunchKernel(ctx->user_reward_traj, dim3((rows + 127) / 128), dim3(128), args, 0, st));
  ctx->launches++;
  return BBMPC_OK;
}

int launch_user_reward_rows(bbmpc_ctx* ctx, const float* s, const float* a, const float* s2, float* out, int B, cudaStream_t st) {
  if (!ctx->user_reward_rows) return fail(ctx, BBMPC_ESTATE, "no user reward is compiled");
  if (ctx->user_reward_dS != ctx->model.dS || ctx->user_reward_dU != ctx->model.dU)
    return fail(ctx, BBMPC_ESTATE, "the user reward was compiled for dS=%d dU=%d", ctx->user_reward_dS, ctx->user_reward_dU);
  void* args[] = {&s, &a, &s2, &out, &B};
  BB_CUDA(ctx, cudaLaunchKernel(ctx->user_reward_rows, dim3((B + 127) / 128), dim3(128), args, 0, st));
  ctx->launches++;
  return BBMPC_OK;
}

}  // namespace bbmpc

extern "C" int bbmpc_reward_set_nvrtc(bbmpc_ctx* ctx, const char* cuda_source) {
  if (!ctx) return BBMPC_EINVAL;
  if (!cuda_source) return bbmpc::fail(ctx, BBMPC_EINVAL, "NULL source");
  if (cudaSetDevice(ctx->device) != cudaSuccess) return bbmpc::fail(ctx, BBMPC_ECUDA, "cudaSetDevice failed");
  if (ctx->user_reward_src == cuda_source && ctx->user_reward_traj && ctx->user_reward_dS == ctx->model.dS &&
      ctx->user_reward_dU == ctx->model.dU) {
    if (ctx->reward_id != BBMPC_REWARD_USER) ctx->epoch++;
    ctx->reward_id = BBMPC_REWARD_USER;
    return BBMPC_OK;   // already compiled for this context
  }
  if (int rc = bbmpc::user_reward_compile(ctx, cuda_source)) return rc;
  ctx->user_reward_src = cuda_source;
  ctx->reward_id = BBMPC_REWARD_USER;
  ctx->epoch++;
  return BBMPC_OK;
}
