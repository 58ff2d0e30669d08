// state.h -- the opaque handle types behind include/vbnn.h.
#pragma once
#include <vector>

#include "common.cuh"
#include "gemm.h"
#include "kernels.h"

struct vbnn_ctx {
  int device = 0;
  cudaStream_t stream = nullptr;
  bool own_stream = false;
  cudaStream_t copy_stream = nullptr;
  uint64_t seed = 0;
  uint32_t* d_step = nullptr;        // Philox "minibatch" counter, lives on the device (graph-safe)
  double* d_partials = nullptr;      // kMaxPartials * kStatSlots doubles of reduction scratch
  double* h_partials = nullptr;      // pinned mirror
  float* h_scalars = nullptr;        // pinned scratch for scalar read-backs
  int next_layer_id = 0;
  long long launches = 0;
  // per-launch device timing of the tensor-core GEMM (bench.py roofline)
  bool profiling = false;
  struct ProfRec { cudaEvent_t a, b; int cls; double flops; };
  std::vector<ProfRec> prof_recs;
  std::vector<cudaEvent_t> prof_pool;
  double prof_ms[8] = {0}; double prof_flops[8] = {0}; long long prof_n[8] = {0};
  // data parallel
  void* nccl_comm = nullptr;
  int rank = 0, nranks = 1;
  cudaStream_t comm_stream = nullptr;   // per-layer gradient allreduce overlaps the rest of backward
};

struct vbnn_layer {
  vbnn_ctx* ctx = nullptr;
  int I = 0, O = 0, kind = VBNN_KIND_VB, id = 0;
  vbnn_opts opts;
  int ldI = 0, ldO = 0;              // padded leading dims (multiples of 8) of operand buffers
  int S_alloc = 1;                   // weight samples held at once
  bool owned_by_mlp = false;
  bool grads_external = false;       // gW/gS/gb live in the mlp's allreduce arena
  // fp32 master state, dense [O x I] / [O]
  float *means = nullptr, *lvars = nullptr, *bias = nullptr, *weight = nullptr;
  float *gW = nullptr, *gS = nullptr, *gb = nullptr;
  float *m_mu = nullptr, *v_mu = nullptr, *m_var = nullptr, *v_var = nullptr;
  float *eps = nullptr, *stdv = nullptr, *mu_sqe = nullptr, *s2_f32 = nullptr;
  float* var_hat_dev = nullptr;
  double* prior_partials = nullptr;  // 2 x kMaxPartials: per-block sums written by the last update (ping-pong:
                                     // an update reads one half while its blocks write the other)
  // half (t & 1) holds the sums of the CURRENT parameters, t = the layer's device step counter
  int* t_dev = nullptr;              // meanState.t == varState.t == biasState.evalCounter
  // bf16 tensor-core operand copies [.. x ldI]
  bf16 *w_bf16 = nullptr, *mu_bf16 = nullptr, *s2_bf16 = nullptr;
  // layer-API scratch (grown on demand)
  int cap_N = 0;
  void *xs = nullptr, *xs2 = nullptr, *gs_ = nullptr, *hs = nullptr, *R = nullptr;
  float* zeta_keep = nullptr;
  int cur_sample = 0;
  bool eps_injected = false;
  bool map_mode = false;
  bool prior_valid = false;
};

struct vbnn_mlp {
  vbnn_ctx* ctx = nullptr;
  std::vector<int> sizes;
  std::vector<vbnn_layer*> layers;   // hidden VB layers..., then the output layer
  int vb_output = 0, max_batch = 0, Z = 1;
  vbnn_opts opts;
  bool bf16 = false, lrt = false;
  std::vector<int> ld;               // padded width of sizes[k]
  // activations, element type float (FP32) or bf16 (BF16); index k = features sizes[k]
  std::vector<void*> act, act2;      // act[0] = staged input [N x ld0]; act[k>0] = [Z x N x ld_k]
  std::vector<void*> R, G, H;        // per layer output k+1: [Z x N x ld_{k+1}]
  float* logits = nullptr; int ld_logits = 0;
  float* logp = nullptr;
  float* targets = nullptr;          // staged targets [N]
  float* xstage[2] = {nullptr, nullptr};   // fp32 H2D staging for the host-buffer API
  float* tstage[2] = {nullptr, nullptr};
  float* result_acc = nullptr;       // [2*Z] loss sums / correct counts
  float* result = nullptr;           // [2] {error, accuracy}
  float* grad_arena = nullptr; size_t grad_count = 0;
  std::vector<size_t> grad_off, grad_len;       // per-layer slice {gW, gS, gb} of the arena
  std::vector<cudaEvent_t> ev_bwd, ev_red;      // dW of layer j done / its allreduce done
  int** t_list_dev = nullptr; int n_t = 0;
  // CUDA graph of one step, keyed by N
  cudaGraphExec_t graph = nullptr; int graph_N = -1; int eager_steps = 0; bool use_graph = true;
  long long graph_launches = 0;
  // host-buffer pipeline
  struct Slot { cudaEvent_t copied, consumed, done; float* h_result; bool busy; int N; };
  Slot slots[2];
  int submit_idx = 0, collect_idx = 0, inflight = 0;
  bool pipeline_ready = false;
  int last_N = 0;
};

namespace vbnn {
int layer_create_internal(vbnn_ctx* ctx, int I, int O, int kind, const vbnn_opts* opts, int S_alloc,
                          float* gW, float* gS, float* gb, int id, vbnn_layer** out);
int layer_update_internal(vbnn_layer* L, vbnn_stats* stats, bool bump_t);
int layer_refresh_copies(vbnn_layer* L);
int layer_compute_prior_internal(vbnn_layer* L);
int layer_refresh_prior_partials(vbnn_layer* L);
PhiloxStream layer_stream(const vbnn_layer* L, uint32_t kind, int sample);
int comm_allreduce_internal(vbnn_ctx* ctx, float* buf, size_t count, cudaStream_t st);
// tensor-core GEMM launch with optional event bracketing (ctx->profiling)
int tc_gemm(vbnn_ctx* ctx, int mode, const TcGemmArgs& g, const EpiParams& p);
int prof_collect(vbnn_ctx* ctx);
}  // namespace vbnn
