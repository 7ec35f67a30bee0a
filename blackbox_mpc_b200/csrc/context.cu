// context.cu — context lifetime, model staging (fp32 image + bf16 hi/lo tensor-core image),
// evaluator-level C-ABI entry points.
#include <cstdarg>
#include <cstring>
#include <cstdlib>
#include <cuda_bf16.h>
#include "common.cuh"
#include "device_fns.cuh"

namespace bbmpc {

static std::string g_create_err;

int fail(bbmpc_ctx* ctx, int code, const char* fmt, ...) {
  char buf[1024];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof buf, fmt, ap);
  va_end(ap);
  if (ctx) ctx->err = buf; else g_create_err = buf;
  return code;
}

static inline int round_up(int x, int m) { return (x + m - 1) / m * m; }

// ---------------------------------------------------------------------------- staging kernels
__global__ void copy_pad_kernel(const float* __restrict__ src, float* __restrict__ dst, int K, int N,
                                int ldw) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= K * ldw) return;
  const int k = i / ldw, n = i % ldw;
  dst[i] = n < N ? src[static_cast<size_t>(k) * N + n] : 0.0f;
}

// Tensor-core image of one (member, layer): weight rows (times `scale`) split hi + lo; three rows hold the bias as
// bf16 terms (b = b1 + b2 + b3 to 24 bits) that meet ones-columns of the A operand; the rest is
// zero.  Image row order: hidden layers [w_0..w_{K-1} | bias x3]; layer 0 (du_slots > 0) is
// [action rows padded to du_slots | state rows | bias x3] so that the epilogue can assemble its
// input with compile-time register indices.
__global__ void pack_tc_kernel(const float* __restrict__ W, const float* __restrict__ b,
                               uint8_t* __restrict__ img, int K, int N, int ldw, int Kpad, int Npad, int nsplit,
                               int du_slots, int dS, int dU, float scale) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= Kpad * Npad) return;
  const int k = i / Npad, n = i % Npad;
  int src = -1, bias_term = -1;
  if (du_slots > 0) {
    if (k < du_slots) src = (k < dU) ? dS + k : -1;
    else if (k - du_slots < dS) src = k - du_slots;
    else if (k - du_slots - dS < BIAS_COLS) bias_term = k - du_slots - dS;
  } else {
    if (k < K) src = k;
    else if (k - K < BIAS_COLS) bias_term = k - K;
  }
  __nv_bfloat16 hi = __float2bfloat16(0.0f), lo = hi;
  if (n < N) {
    if (src >= 0) {
      const float v = __fmul_rn(W[static_cast<size_t>(src) * ldw + n], scale);
      hi = __float2bfloat16_rn(v);
      lo = __float2bfloat16_rn(v - __bfloat162float(hi));
    } else if (bias_term >= 0) {
      float r = __fmul_rn(b[n], scale);
      __nv_bfloat16 t = __float2bfloat16_rn(r);
      for (int j = 0; j < bias_term; ++j) { r -= __bfloat162float(t); t = __float2bfloat16_rn(r); }
      hi = t;
    }
  }
  const int chunk = k >> 4, kk = k & 15;
  const bool second = nsplit > 0 && n >= nsplit;
  const int Nh = nsplit > 0 ? (second ? Npad - nsplit : nsplit) : Npad;
  const int nh = second ? n - nsplit : n;
  const size_t half_base = second ? static_cast<size_t>(Kpad / 16) * nsplit * 64 : 0;
  const size_t off = half_base + static_cast<size_t>(chunk) * Nh * 64 + (static_cast<size_t>(kk >> 3) * Nh + nh) * 16 + (kk & 7) * 2;
  *reinterpret_cast<__nv_bfloat16*>(img + off) = hi;
  *reinterpret_cast<__nv_bfloat16*>(img + off + static_cast<size_t>(Nh) * 32) = lo;
}

__global__ void norm_prepare_kernel(const float* __restrict__ std_in, float* __restrict__ den, int n) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) den[i] = __fadd_rn(std_in[i], 1e-7f);
}

// ---------------------------------------------------------------------------- tensor-core eligibility
bool tc_supported(const ModelHost& m, std::string* why) {
  auto no = [&](const char* s) { if (why) *why = s; return false; };
  if (m.dyn_id != BBMPC_DYN_MLP) return no("analytical dynamics has no GEMM");
  const MlpDev& p = m.mlp;
  if (m.dS > 32 || m.dU > 16) return no("dS > 32 or dU > 16");
  for (int l = 0; l < p.n_layers; ++l) {
    const LayerDev& L = p.layer[l];
    if (L.Npad > 256 || L.Kpad > 256) return no("layer wider than 253");
  }
  if (!tc_column_map(p, nullptr, nullptr, nullptr))
    return no("TMEM budget exceeded (hidden width + 3 must pad to <= 224 columns)");
  if (p.layer[p.n_layers - 1].N != m.dS) return no("output width != dS");
  if (p.n_members > 1 && p.layer[p.n_layers - 1].act != BBMPC_ACT_NONE)
    return no("ensemble with a non-linear output layer");
  return true;
}

int resolve_precision(const bbmpc_ctx* ctx) {
  if (ctx->model.dyn_id != BBMPC_DYN_MLP) return BBMPC_PREC_FP32;
  if (ctx->prec == BBMPC_PREC_AUTO) return ctx->model.tc_ok ? BBMPC_PREC_BF16X3 : BBMPC_PREC_FP32;
  return ctx->prec;
}

static void free_model(ModelHost& m) {
  cudaFree(m.w32_buf); cudaFree(m.wimg_buf); cudaFree(m.chunk_table_buf); cudaFree(m.jobs_buf);
  cudaFree(m.solo_table_buf); cudaFree(m.solo_jobs_buf); m.solo_table_buf = nullptr; m.solo_jobs_buf = nullptr;
  m.w32_buf = nullptr; m.wimg_buf = nullptr; m.chunk_table_buf = nullptr; m.jobs_buf = nullptr;
  m.mlp = MlpDev{};
  m.tc_ok = false;
  m.set = false;
}

int rollout_dispatch(bbmpc_ctx* ctx, const float* states, const float* actions, float* returns,
                     const float* penalty, int rows, int A, int H, cudaStream_t st) {
  if (!ctx->model.set) return fail(ctx, BBMPC_ESTATE, "rollout before a dynamics model was set");
  if (!ctx->reward_id) return fail(ctx, BBMPC_ESTATE, "rollout before a reward function was set");
  const int prec = resolve_precision(ctx);
  if (prec != BBMPC_PREC_FP32 && !ctx->model.tc_ok)
    return fail(ctx, BBMPC_EINVAL, "tensor-core precision requested but the model does not fit it: %s",
                ctx->model.tc_why.c_str());
  cudaEvent_t e0 = nullptr, e1 = nullptr;
  if (ctx->prof_on) {
    if (ctx->prof_ev.size() < 2 * (ctx->prof_n + 1)) {
      cudaEvent_t a, b;
      BB_CUDA(ctx, cudaEventCreate(&a));
      BB_CUDA(ctx, cudaEventCreate(&b));
      ctx->prof_ev.push_back(a); ctx->prof_ev.push_back(b);
    }
    e0 = ctx->prof_ev[2 * ctx->prof_n]; e1 = ctx->prof_ev[2 * ctx->prof_n + 1];
    BB_CUDA(ctx, cudaEventRecord(e0, st));
  }
  int rc;
  // user reward: the kernels dump the visited states (their own reward is off), a JIT-compiled kernel sums the reward
  const bool user_reward = ctx->reward_id == BBMPC_REWARD_USER;
  ctx->traj_cur = nullptr;
  if (user_reward && H > 0)
    if (int rc2 = user_reward_traj_buffer(ctx, rows, H, st, &ctx->traj_cur)) return rc2;
  if (prec == BBMPC_PREC_FP32) {
    rc = launch_rollout_simt(ctx, states, actions, returns, penalty, rows, A, H, 0, st);
    ctx->last_rollout_kernel = "rollout_simt_kernel";
  } else {
    // pipelined kernel (two jobs in flight per CTA) when the model fits its TMEM / shared-memory budget;
    // BBMPC_TC_PIPE=0 keeps the one-tile-per-CTA kernel (A/B measurements)
    const int passes = prec == BBMPC_PREC_BF16 ? 1 : 3;
    const char* pipe_env = getenv("BBMPC_TC_PIPE");
    rc = -100;
    if (H > 0 && !(pipe_env && pipe_env[0] == '0')) rc = launch_rollout_pipe(ctx, states, actions, returns, penalty, rows, A, H, passes, st);
    ctx->last_rollout_kernel = "rollout_pipe_kernel";
    if (rc == -100) { rc = launch_rollout_tc(ctx, states, actions, returns, penalty, rows, A, H, passes, st); ctx->last_rollout_kernel = "rollout_tc_kernel"; }
  }
  if (rc == BBMPC_OK && user_reward)
    rc = launch_user_reward_traj(ctx, ctx->traj_cur, states, actions, penalty, returns, rows, A, H, st);
  ctx->traj_cur = nullptr;
  if (rc == BBMPC_OK && e1) {
    BB_CUDA(ctx, cudaEventRecord(e1, st));
    ctx->prof_n++;
  }
  return rc;
}

}  // namespace bbmpc

using namespace bbmpc;

extern "C" {

int bbmpc_version(void) { return 200; }
int bbmpc_abi_config_size(void) { return static_cast<int>(sizeof(bbmpc_opt_config)); }

const char* bbmpc_last_error(const bbmpc_ctx* ctx) { return ctx ? ctx->err.c_str() : g_create_err.c_str(); }

int bbmpc_ctx_create(int device, uint64_t seed, bbmpc_ctx** out) {
  if (!out) return fail(nullptr, BBMPC_EINVAL, "out is NULL");
  *out = nullptr;
  int n = 0;
  cudaError_t e = cudaGetDeviceCount(&n);
  if (e != cudaSuccess || n == 0)
    return fail(nullptr, BBMPC_ECUDA, "no CUDA device available (%s); libbbmpc has no CPU fallback",
                e != cudaSuccess ? cudaGetErrorString(e) : "device count is 0");
  if (device < 0 || device >= n) return fail(nullptr, BBMPC_EINVAL, "device %d out of range [0,%d)", device, n);
  cudaDeviceProp prop;
  if ((e = cudaGetDeviceProperties(&prop, device)) != cudaSuccess)
    return fail(nullptr, BBMPC_ECUDA, "cudaGetDeviceProperties: %s", cudaGetErrorString(e));
  if (prop.major != 10)
    return fail(nullptr, BBMPC_ECUDA, "device %d is sm_%d%d; libbbmpc is built for sm_100a only", device,
                prop.major, prop.minor);
  if ((e = cudaSetDevice(device)) != cudaSuccess)
    return fail(nullptr, BBMPC_ECUDA, "cudaSetDevice: %s", cudaGetErrorString(e));
  bbmpc_ctx* c = new bbmpc_ctx();
  c->device = device;
  c->seed = seed;
  c->sm_count = prop.multiProcessorCount;
  *out = c;
  return BBMPC_OK;
}

void bbmpc_ctx_destroy(bbmpc_ctx* ctx) {
  if (!ctx) return;
  cudaSetDevice(ctx->device);
  free_model(ctx->model);
  cudaFree(ctx->model.norm_buf);
  cudaFree(ctx->step_scratch); cudaFree(ctx->step_counters);
  cudaFree(ctx->tc_xchg); cudaFree(ctx->tc_flags); cudaFree(ctx->pipe_park);
  user_reward_free(ctx);
  if (ctx->dbg_host) cudaFreeHost(ctx->dbg_host);
  for (cudaEvent_t e : ctx->prof_ev) cudaEventDestroy(e);
  delete ctx;
}

int bbmpc_set_precision(bbmpc_ctx* ctx, int prec) {
  if (!ctx) return BBMPC_EINVAL;
  if (prec < BBMPC_PREC_AUTO || prec > BBMPC_PREC_BF16) return fail(ctx, BBMPC_EINVAL, "unknown precision %d", prec);
  if (ctx->prec != prec) ctx->epoch++;
  ctx->prec = prec;
  return BBMPC_OK;
}

int bbmpc_get_effective_precision(const bbmpc_ctx* ctx) { return ctx ? resolve_precision(ctx) : BBMPC_EINVAL; }

uint64_t bbmpc_launch_count(const bbmpc_ctx* ctx) { return ctx ? ctx->launches : 0; }

const char* bbmpc_last_rollout_kernel(const bbmpc_ctx* ctx) { return ctx ? ctx->last_rollout_kernel.c_str() : ""; }

int bbmpc_profile_enable(bbmpc_ctx* ctx, int on) {
  if (!ctx) return BBMPC_EINVAL;
  ctx->prof_on = on != 0;
  ctx->prof_n = 0;
  return BBMPC_OK;
}

int bbmpc_profile_read(bbmpc_ctx* ctx, double* ms_total, int64_t* n_launches) {
  if (!ctx) return BBMPC_EINVAL;
  BB_CUDA(ctx, cudaSetDevice(ctx->device));
  double total = 0.0;
  for (size_t i = 0; i < ctx->prof_n; ++i) {
    BB_CUDA(ctx, cudaEventSynchronize(ctx->prof_ev[2 * i + 1]));
    float ms = 0.0f;
    BB_CUDA(ctx, cudaEventElapsedTime(&ms, ctx->prof_ev[2 * i], ctx->prof_ev[2 * i + 1]));
    total += ms;
  }
  if (ms_total) *ms_total = total;
  if (n_launches) *n_launches = static_cast<int64_t>(ctx->prof_n);
  ctx->prof_n = 0;
  return BBMPC_OK;
}

int bbmpc_model_set_mlp(bbmpc_ctx* ctx, int n_members, int n_layers, const int* dims,
                        const float* const* W, const float* const* b, const int* act_ids, void* stream) {
  if (!ctx) return BBMPC_EINVAL;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (n_members < 1 || n_members > MAX_MEMBERS) return fail(ctx, BBMPC_EINVAL, "n_members %d not in [1,%d]", n_members, MAX_MEMBERS);
  if (n_layers < 1 || n_layers > MAX_LAYERS) return fail(ctx, BBMPC_EINVAL, "n_layers %d not in [1,%d]", n_layers, MAX_LAYERS);
  if (!dims || !W || !b || !act_ids) return fail(ctx, BBMPC_EINVAL, "NULL argument");
  for (int l = 0; l <= n_layers; ++l)
    if (dims[l] < 1 || dims[l] > 4096) return fail(ctx, BBMPC_EINVAL, "layer width %d unsupported", dims[l]);
  BB_CUDA(ctx, cudaSetDevice(ctx->device));
  ModelHost& m = ctx->model;
  BB_CUDA(ctx, cudaStreamSynchronize(st));  // nothing in flight may still read the old image
  ctx->epoch++;
  free_model(m);
  m.dyn_id = BBMPC_DYN_MLP;
  MlpDev& p = m.mlp;
  p.n_members = n_members;
  p.n_layers = n_layers;
  // dS/dU are fixed by set_norm / opt_create when called first; else infer them from the dims
  if (m.dS == 0) { m.dS = dims[n_layers]; m.dU = dims[0] - dims[n_layers]; }
  if (m.dS + m.dU != dims[0] || m.dS != dims[n_layers])
    return fail(ctx, BBMPC_EINVAL, "MLP dims [%d -> %d] do not match dS=%d dU=%d", dims[0], dims[n_layers], m.dS, m.dU);
  if (m.dS > MAX_DS || m.dU > MAX_DU || m.dU < 1) return fail(ctx, BBMPC_EINVAL, "dS=%d dU=%d unsupported (max %d/%d)", m.dS, m.dU, MAX_DS, MAX_DU);
  int64_t w_floats = 0, img_bytes = 0;
  p.max_width = dims[0];
  for (int l = 0; l < n_layers; ++l) {
    LayerDev& L = p.layer[l];
    L.K = dims[l]; L.N = dims[l + 1]; L.ldw = round_up(L.N, 16); L.act = act_ids[l];
    if (L.act < BBMPC_ACT_NONE || L.act > BBMPC_ACT_SIGMOID) return fail(ctx, BBMPC_EINVAL, "unknown activation id %d", L.act);
    L.Kpad = round_up((l == 0 ? L.K - m.dU + tc_du_slots(m.dU) : L.K) + BIAS_COLS, 16);
    L.Npad = round_up(L.N, 16);
    // Optional (BBMPC_NSPLIT=1): hidden layers wide enough are cut into two column ranges at a multiple of 32
    // (whole epilogue units).  Measured slower than the unsplit pipeline on B200 (r1c: 1.90 vs 1.73 ms per
    // C4 rollout): the extra handoffs cost more than the overlap gains while the epilogue is MUFU-bound.
    L.nsplit = (l + 1 < n_layers && L.Npad >= 64 && getenv("BBMPC_NSPLIT")) ? (L.Npad / 32) * 16 : 0;
    if (L.N > p.max_width) p.max_width = L.N;
    L.w_off = w_floats; w_floats += static_cast<int64_t>(L.K) * L.ldw;
    L.b_off = w_floats; w_floats += L.ldw;
    L.img_off = img_bytes; img_bytes += static_cast<int64_t>(L.Kpad / 16) * L.Npad * 64;
  }
  p.w_member_stride = w_floats;
  p.img_member_stride = img_bytes;
  w_floats *= n_members;
  img_bytes *= n_members;
  BB_CUDA(ctx, cudaMalloc(&m.w32_buf, w_floats * sizeof(float)));
  BB_CUDA(ctx, cudaMemsetAsync(m.w32_buf, 0, w_floats * sizeof(float), st));
  for (int l = 0; l < n_layers; ++l) {
    const LayerDev& L = p.layer[l];
    for (int mm = 0; mm < n_members; ++mm) {
      const float* Wsrc = W[mm * n_layers + l];
      const float* bsrc = b[mm * n_layers + l];
      if (!Wsrc || !bsrc) return fail(ctx, BBMPC_EINVAL, "NULL weight pointer (member %d layer %d)", mm, l);
      const int n = L.K * L.ldw;
      copy_pad_kernel<<<(n + 255) / 256, 256, 0, st>>>(Wsrc, m.w32_buf + L.w_off + mm * p.w_member_stride, L.K, L.N, L.ldw);
      BB_LAUNCH_CHECK(ctx);
      BB_CUDA(ctx, cudaMemcpyAsync(m.w32_buf + L.b_off + mm * p.w_member_stride, bsrc, L.N * sizeof(float), cudaMemcpyDeviceToDevice, st));
    }
  }
  p.w32 = m.w32_buf;
  m.set = true;
  // tensor-core image
  m.tc_ok = tc_supported(m, &m.tc_why);
  if (m.tc_ok) {
    BB_CUDA(ctx, cudaMalloc(&m.wimg_buf, img_bytes));
    std::vector<uint2> table;
    for (int mm = 0; mm < n_members; ++mm)
      for (int l = 0; l < n_layers; ++l) {
        const LayerDev& L = p.layer[l];
        const int n = L.Kpad * L.Npad;
        const int64_t io = L.img_off + mm * p.img_member_stride;
        pack_tc_kernel<<<(n + 255) / 256, 256, 0, st>>>(m.w32_buf + L.w_off + mm * p.w_member_stride,
                                                       m.w32_buf + L.b_off + mm * p.w_member_stride,
                                                       m.wimg_buf + io, L.K, L.N, L.ldw, L.Kpad, L.Npad, L.nsplit,
                                                       l == 0 ? tc_du_slots(m.dU) : 0, m.dS, m.dU,
                                                       // hidden tanh layers deliver 2 log2(e) x: the epilogue's tanh starts at ex2
                                                       (l + 1 < n_layers && L.act == BBMPC_ACT_TANH) ? 2.8853900817779268f : 1.0f);
        BB_LAUNCH_CHECK(ctx);
      }
    // chunk table in the MMA issue order of rollout_tc_kernel (run_layer): L0(0); per member: hidden
    // layers 1..nL-2, L0(next member), output layer.
    std::vector<TcJob> jobs;
    int buf_w = 0, col_x = 0, col_dout = 0;
    tc_column_map(p, &buf_w, &col_x, &col_dout);
    int stage_bytes = 0;
    // One job per (member, layer, column half), in MMA issue order.
    bool solo = false;
    auto push_layer = [&](int mm, int l) {
      const LayerDev& L = p.layer[l];
      const int nch = L.Kpad / 16;
      const int nunits = (nch + 1) / 2;
      const bool last = (l == n_layers - 1);
      const int idx = mm * (n_layers - 1) + l;   // running index of hidden accumulators: TMEM buffer = idx & 1
      const int n_halves = L.nsplit > 0 ? 2 : 1;
      for (int h = 0; h < n_halves; ++h) {
        const int n0 = h ? L.nsplit : 0;
        const int Nh = L.nsplit > 0 ? (h ? L.Npad - L.nsplit : L.nsplit) : L.Npad;
        const int chunk_bytes = Nh * 64;
        const int64_t io = L.img_off + mm * p.img_member_stride + (h ? static_cast<int64_t>(nch) * L.nsplit * 64 : 0);
        // groups of whole UNITS (pairs of K-chunks): one ring stage / bulk copy / barrier pair each
        int gmax = TC_GROUP_CAP_BYTES / (2 * chunk_bytes);
        if (gmax < 1) gmax = 1;
        if (gmax > 15) gmax = 15;
        int ngroups = (nunits + gmax - 1) / gmax;
        if (ngroups > TC_MAX_GROUPS) ngroups = TC_MAX_GROUPS;
        uint32_t gsz = 0;
        int u0 = 0;
        for (int g = 0; g < ngroups; ++g) {
          const int nu = (nunits - u0 + (ngroups - g) - 1) / (ngroups - g);   // balanced split
          const int c0 = 2 * u0, c1 = (2 * (u0 + nu) < nch) ? 2 * (u0 + nu) : nch;
          table.push_back(make_uint2(static_cast<uint32_t>(io + static_cast<int64_t>(c0) * chunk_bytes),
                                     static_cast<uint32_t>((c1 - c0) * chunk_bytes)));
          if ((c1 - c0) * chunk_bytes > stage_bytes) stage_bytes = (c1 - c0) * chunk_bytes;
          gsz |= static_cast<uint32_t>(nu) << (4 * g);
          u0 += nu;
        }
        TcJob j{};
        j.d_col = (last ? col_dout : ((idx & 1) ? 256 : 0)) + n0;
        j.a_col = l == 0 ? col_x : (((idx - 1) & 1) ? 256 : 0);
        j.idesc = tc_idesc(Nh);
        j.nchunks = nch;
        const uint32_t kstep = static_cast<uint32_t>(Nh) * 16;  // bytes between the two 8-wide K slabs of a chunk
        j.desc_lo_base = ((kstep >> 4) & 0x3FFF) << 16;
        j.lo_off16 = (2 * kstep) >> 4;
        j.chunk16 = static_cast<uint32_t>(chunk_bytes) >> 4;
        j.ngroups = ngroups;
        j.gsz = gsz;
        uint32_t commit;
        if (last) commit = (solo || mm == n_members - 1) ? TCJ_COMMIT_DOUT : 0u;
        else if (l == 0) commit = h ? TCJ_COMMIT_D0H1 : TCJ_COMMIT_D0H0;
        else commit = h ? TCJ_COMMIT_DH1 : TCJ_COMMIT_DH0;
        j.flags = (l == 0 && mm == 0 && h == 0 ? TCJ_WAIT_X : 0u) | (l > 0 && h == 0 ? TCJ_FROM_EPI : 0u) |
                  (l > 0 && h == n_halves - 1 ? TCJ_ROUND_END : 0u) | (last && mm > 0 && !solo ? TCJ_ACC_FIRST : 0u) | commit;
        jobs.push_back(j);
      }
    };
    if (n_layers == 1) {
      for (int mm = 0; mm < n_members; ++mm) push_layer(mm, 0);
    } else {
      const bool early = getenv("BBMPC_NO_EARLY_L0") == nullptr;
      p.early_l0 = early ? 1 : 0;
      if (early) push_layer(0, 0);
      for (int mm = 0; mm < n_members; ++mm) {
        if (!early) push_layer(mm, 0);
        for (int l = 1; l + 1 < n_layers; ++l) push_layer(mm, l);
        if (early && mm + 1 < n_members) push_layer(mm + 1, 0);
        push_layer(mm, n_layers - 1);
      }
    }
    // solo tables (member 0 only, plain layer order)
    std::vector<uint2> table_all = table;
    std::vector<TcJob> jobs_all = jobs;
    const int stage_all = stage_bytes;
    table.clear(); jobs.clear(); solo = true;
    for (int l = 0; l < n_layers; ++l) push_layer(0, l);
    std::vector<uint2> table_solo = table;
    std::vector<TcJob> jobs_solo = jobs;
    table = table_all; jobs = jobs_all;
    if (stage_all > stage_bytes) stage_bytes = stage_all;
    BB_CUDA(ctx, cudaMalloc(&m.solo_table_buf, table_solo.size() * sizeof(uint2)));
    BB_CUDA(ctx, cudaMemcpyAsync(m.solo_table_buf, table_solo.data(), table_solo.size() * sizeof(uint2), cudaMemcpyHostToDevice, st));
    BB_CUDA(ctx, cudaMalloc(&m.solo_jobs_buf, jobs_solo.size() * sizeof(TcJob)));
    BB_CUDA(ctx, cudaMemcpyAsync(m.solo_jobs_buf, jobs_solo.data(), jobs_solo.size() * sizeof(TcJob), cudaMemcpyHostToDevice, st));
    p.solo_table = m.solo_table_buf; p.solo_groups_per_step = static_cast<int>(table_solo.size());
    p.solo_jobs = m.solo_jobs_buf; p.solo_jobs_per_step = static_cast<int>(jobs_solo.size());
    BB_CUDA(ctx, cudaMalloc(&m.chunk_table_buf, table.size() * sizeof(uint2)));
    BB_CUDA(ctx, cudaMemcpyAsync(m.chunk_table_buf, table.data(), table.size() * sizeof(uint2), cudaMemcpyHostToDevice, st));
    BB_CUDA(ctx, cudaMalloc(&m.jobs_buf, jobs.size() * sizeof(TcJob)));
    BB_CUDA(ctx, cudaMemcpyAsync(m.jobs_buf, jobs.data(), jobs.size() * sizeof(TcJob), cudaMemcpyHostToDevice, st));
    BB_CUDA(ctx, cudaStreamSynchronize(st));  // `table` / `jobs` are stack-lifetime host buffers
    p.jobs = m.jobs_buf;
    p.jobs_per_step = static_cast<int>(jobs.size());
    p.wimg = m.wimg_buf;
    p.chunk_table = m.chunk_table_buf;
    p.chunks_per_step = static_cast<int>(table.size());
    p.stage_bytes = stage_bytes;
  }
  return BBMPC_OK;
}

int bbmpc_model_set_norm(bbmpc_ctx* ctx, int dS, int dU, const float* mean_s, const float* std_s,
                         const float* mean_a, const float* std_a, const float* mean_t, const float* std_t,
                         void* stream) {
  if (!ctx) return BBMPC_EINVAL;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (dS < 1 || dS > MAX_DS || dU < 1 || dU > MAX_DU) return fail(ctx, BBMPC_EINVAL, "dS=%d dU=%d unsupported", dS, dU);
  BB_CUDA(ctx, cudaSetDevice(ctx->device));
  ModelHost& m = ctx->model;
  if (m.set && m.dyn_id == BBMPC_DYN_MLP && (m.dS != dS || m.dU != dU))
    return fail(ctx, BBMPC_EINVAL, "norm dims dS=%d dU=%d differ from the model's dS=%d dU=%d", dS, dU, m.dS, m.dU);
  ctx->epoch++;
  m.dS = dS; m.dU = dU;
  const int n_null = !mean_s + !std_s + !mean_a + !std_a + !mean_t + !std_t;
  if (n_null == 6) { m.norm = NormDev{}; return BBMPC_OK; }
  if (n_null != 0) return fail(ctx, BBMPC_EINVAL, "normalisation statistics must be all set or all NULL");
  if (!m.norm_buf) BB_CUDA(ctx, cudaMalloc(&m.norm_buf, (4 * MAX_DS + 2 * MAX_DU) * sizeof(float)));
  float* base = m.norm_buf;
  float *ms = base, *ds = base + MAX_DS, *mt = base + 2 * MAX_DS, *dt = base + 3 * MAX_DS,
        *ma = base + 4 * MAX_DS, *da = base + 4 * MAX_DS + MAX_DU;
  BB_CUDA(ctx, cudaMemcpyAsync(ms, mean_s, dS * 4, cudaMemcpyDeviceToDevice, st));
  BB_CUDA(ctx, cudaMemcpyAsync(mt, mean_t, dS * 4, cudaMemcpyDeviceToDevice, st));
  BB_CUDA(ctx, cudaMemcpyAsync(ma, mean_a, dU * 4, cudaMemcpyDeviceToDevice, st));
  norm_prepare_kernel<<<1, 64, 0, st>>>(std_s, ds, dS); BB_LAUNCH_CHECK(ctx);
  norm_prepare_kernel<<<1, 64, 0, st>>>(std_t, dt, dS); BB_LAUNCH_CHECK(ctx);
  norm_prepare_kernel<<<1, 64, 0, st>>>(std_a, da, dU); BB_LAUNCH_CHECK(ctx);
  m.norm = NormDev{1, ms, ds, ma, da, mt, dt};
  return BBMPC_OK;
}

int bbmpc_model_set_builtin(bbmpc_ctx* ctx, int dyn_id, int dS, int dU) {
  if (!ctx) return BBMPC_EINVAL;
  if (dyn_id != BBMPC_DYN_PENDULUM) return fail(ctx, BBMPC_EINVAL, "unknown builtin dynamics id %d", dyn_id);
  if (dS != 3 || dU != 1) return fail(ctx, BBMPC_EINVAL, "pendulum model needs dS=3 dU=1 (got %d, %d)", dS, dU);
  BB_CUDA(ctx, cudaSetDevice(ctx->device));
  BB_CUDA(ctx, cudaDeviceSynchronize());
  ctx->epoch++;
  free_model(ctx->model);
  ctx->model.dyn_id = dyn_id;
  ctx->model.dS = dS; ctx->model.dU = dU;
  ctx->model.norm = NormDev{};
  ctx->model.set = true;
  ctx->model.tc_why = "analytical dynamics has no GEMM";
  return BBMPC_OK;
}

int bbmpc_reward_set_builtin(bbmpc_ctx* ctx, int reward_id) {
  if (!ctx) return BBMPC_EINVAL;
  if (reward_id < BBMPC_REWARD_PENDULUM || reward_id > BBMPC_REWARD_PENDULUM_GYM)
    return fail(ctx, BBMPC_EINVAL, "unknown reward id %d", reward_id);
  if (ctx->reward_id != reward_id) ctx->epoch++;
  ctx->reward_id = reward_id;
  return BBMPC_OK;
}

static int check_reward_dims(bbmpc_ctx* ctx) {
  const ModelHost& m = ctx->model;
  if (ctx->reward_id == BBMPC_REWARD_USER && (ctx->user_reward_dS != m.dS || ctx->user_reward_dU != m.dU))
    return fail(ctx, BBMPC_ESTATE, "the user reward was compiled for dS=%d dU=%d, the model has dS=%d dU=%d: set it again",
                ctx->user_reward_dS, ctx->user_reward_dU, m.dS, m.dU);
  if (ctx->reward_id == BBMPC_REWARD_HALFCHEETAH && m.dS < 18)
    return fail(ctx, BBMPC_EINVAL, "HalfCheetah reward reads state[17]; dS=%d", m.dS);
  if ((ctx->reward_id == BBMPC_REWARD_PENDULUM || ctx->reward_id == BBMPC_REWARD_PENDULUM_GYM) && m.dS < 3)
    return fail(ctx, BBMPC_EINVAL, "pendulum reward reads state[0..2]; dS=%d", m.dS);
  return BBMPC_OK;
}

int bbmpc_rollout(bbmpc_ctx* ctx, const float* states, const float* actions, float* returns, int P, int A,
                  int H, void* stream) {
  if (!ctx) return BBMPC_EINVAL;
  if (P < 0 || A < 1 || H < 0) return fail(ctx, BBMPC_EINVAL, "bad shape P=%d A=%d H=%d", P, A, H);
  if (P == 0) return BBMPC_OK;
  if (!states || !actions || !returns) return fail(ctx, BBMPC_EINVAL, "NULL pointer");
  if (int rc = check_reward_dims(ctx)) return rc;
  BB_CUDA(ctx, cudaSetDevice(ctx->device));
  return rollout_dispatch(ctx, states, actions, returns, nullptr, P * A, A, H, static_cast<cudaStream_t>(stream));
}

int bbmpc_predict_next_state(bbmpc_ctx* ctx, const float* s, const float* a, float* out, int B, void* stream) {
  if (!ctx) return BBMPC_EINVAL;
  if (!ctx->model.set) return fail(ctx, BBMPC_ESTATE, "predict_next_state before a dynamics model was set");
  if (B <= 0) return B == 0 ? BBMPC_OK : fail(ctx, BBMPC_EINVAL, "B < 0");
  BB_CUDA(ctx, cudaSetDevice(ctx->device));
  StepIO io{s, a, nullptr, out, nullptr, nullptr, B, 1};
  return launch_step_simt(ctx, io, static_cast<cudaStream_t>(stream));
}

int bbmpc_reward(bbmpc_ctx* ctx, const float* s, const float* a, const float* s2, float* out, int B, void* stream) {
  if (!ctx) return BBMPC_EINVAL;
  if (!ctx->reward_id) return fail(ctx, BBMPC_ESTATE, "reward before a reward function was set");
  if (!ctx->model.dS) return fail(ctx, BBMPC_ESTATE, "reward before dS/dU are known (set a model first)");
  if (int rc = check_reward_dims(ctx)) return rc;
  if (B <= 0) return B == 0 ? BBMPC_OK : fail(ctx, BBMPC_EINVAL, "B < 0");
  BB_CUDA(ctx, cudaSetDevice(ctx->device));
  if (ctx->reward_id == BBMPC_REWARD_USER) return launch_user_reward_rows(ctx, s, a, s2, out, B, static_cast<cudaStream_t>(stream));
  StepIO io{s, a, s2, nullptr, out, nullptr, B, 2};
  return launch_step_simt(ctx, io, static_cast<cudaStream_t>(stream));
}

int bbmpc_dynamics_forward(bbmpc_ctx* ctx, const float* x, float* out, int B, void* stream) {
  if (!ctx) return BBMPC_EINVAL;
  if (!ctx->model.set) return fail(ctx, BBMPC_ESTATE, "dynamics_forward before a dynamics model was set");
  if (B <= 0) return B == 0 ? BBMPC_OK : fail(ctx, BBMPC_EINVAL, "B < 0");
  BB_CUDA(ctx, cudaSetDevice(ctx->device));
  StepIO io{x, nullptr, nullptr, nullptr, nullptr, out, B, 4};
  return launch_step_simt(ctx, io, static_cast<cudaStream_t>(stream));
}

void bbmpc_philox4x32_host(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4]) {
  const Philox4 r = philox4x32_10(Philox4{ctr[0], ctr[1], ctr[2], ctr[3]}, key[0], key[1]);
  out[0] = r.x; out[1] = r.y; out[2] = r.z; out[3] = r.w;
}

}  // extern "C"
