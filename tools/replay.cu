// replay.cu -- a NON-PYTHON host driving libvbnn.so's layer-level C ABI in exactly the order the
// reference's Lua would (no Lua / LuaJIT / Torch7 exists in this image, so
// this C++ program stands in for `th main.lua` on the hot path).
//
// It links only libvbnn.so + the CUDA runtime and replays, for every minibatch of a golden fixture
// (tests/golden/mlp_weight.replay.bin, written by tests/golden/make_golden.py from the oracle):
//
//   mlp:buildModel        mlp.lua:7-60     VBLinear(24,16) ReLU VBLinear(16,12) ReLU Linear(12,5) LogSoftMax
//   getParameters()       mlp.lua:37       weight/bias/gradWeight/gradBias of every module re-flattened into
//                                          ONE storage -> vbnn_layer_bind makes the VB layers follow them
//   net:resetGradients()  main.lua:28      gradParameters:zero() (a foreign memset) + VBLinear:resetAcc
//   for s = 1, S          main.lua:32-37   net:sample() -> vbnn_layer_sample (fixture epsilon injected)
//     net:run()           mlp.lua:76-84    forward, criterion:backward, model:backward, criterion:forward,
//                                          get_accuracy (host loop over rows, utils.lua:11-27)
//   net:update(opt)       main.lua:40      optim.sgd on the output layer (mlp.lua:120-123), VBLinear:update
//
// Everything that is NOT the VB layer -- nn.ReLU, nn.Linear (output layer), nn.LogSoftMax,
// ClassNLLCriterion, optim.sgd, gradParameters:zero() -- is "foreign" work done by this file's own naive
// kernels on the LEGACY DEFAULT stream (stream 0), as cutorch / cunn modules do.  The library must order
// itself with that work: the context is created with VBNN_CTX_STREAM_LEGACY_DEFAULT (or, second mode,
// VBNN_CTX_STREAM_PRIVATE_BLOCKING).  No explicit synchronisation is issued between foreign kernels and
// library calls; the result must still match the oracle's fixture.
//
//   usage: replay <fixture.bin> [legacy|blocking|nonblocking]
//   ("nonblocking" = the round-1 behaviour, a private cudaStreamNonBlocking stream: shown to be unordered
//    with stream 0 -- the program inserts no syncs, so it is EXPECTED to be able to fail; used only to
//    demonstrate why the flag exists, never by the tests as a pass criterion.)
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <map>
#include <string>
#include <vector>

#include "../include/vbnn.h"

#define CK(x)                                                                                   \
  do {                                                                                          \
    cudaError_t e_ = (x);                                                                       \
    if (e_ != cudaSuccess) { fprintf(stderr, "CUDA %s: %s (%s:%d)\n", #x, cudaGetErrorString(e_), __FILE__, __LINE__); exit(2); } \
  } while (0)
#define VB(x)                                                                                   \
  do {                                                                                          \
    int r_ = (x);                                                                               \
    if (r_ != VBNN_OK) { fprintf(stderr, "libvbnn %s -> %d: %s (%s:%d)\n", #x, r_, vbnn_last_error(), __FILE__, __LINE__); exit(3); } \
  } while (0)

// ------------------------------------------------------------------ foreign kernels (stream 0) ------
__global__ void k_relu(const float* x, float* y, int n) {                 // nn.ReLU:updateOutput
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) y[i] = x[i] > 0.f ? x[i] : 0.f;
}
__global__ void k_relu_bwd(const float* y, const float* g, float* gi, int n) {   // nn.ReLU:updateGradInput
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) gi[i] = y[i] > 0.f ? g[i] : 0.f;
}
__global__ void k_linear_fwd(const float* X, const float* W, const float* b, float* Y, int N, int I, int O) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= N * O) return;
  int n = i / O, o = i % O;
  float a = b[o];
  for (int k = 0; k < I; ++k) a += X[n * I + k] * W[o * I + k];
  Y[i] = a;
}
__global__ void k_linear_bwd_data(const float* G, const float* W, float* dX, int N, int I, int O) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= N * I) return;
  int n = i / I, k = i % I;
  float a = 0.f;
  for (int o = 0; o < O; ++o) a += G[n * O + o] * W[o * I + k];
  dX[i] = a;
}
__global__ void k_linear_acc(const float* X, const float* G, float* gW, float* gb, int N, int I, int O) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < O * I) {
    int o = i / I, k = i % I;
    float a = 0.f;
    for (int n = 0; n < N; ++n) a += G[n * O + o] * X[n * I + k];
    gW[i] += a;
  }
  if (i < O) {
    float a = 0.f;
    for (int n = 0; n < N; ++n) a += G[n * O + i];
    gb[i] += a;
  }
}
__global__ void k_logsoftmax(const float* x, float* lp, int N, int C) {   // nn.LogSoftMax
  int n = blockIdx.x * blockDim.x + threadIdx.x;
  if (n >= N) return;
  float mx = -3e38f;
  for (int c = 0; c < C; ++c) mx = fmaxf(mx, x[n * C + c]);
  float s = 0.f;
  for (int c = 0; c < C; ++c) s += expf(x[n * C + c] - mx);
  float lse = mx + logf(s);
  for (int c = 0; c < C; ++c) lp[n * C + c] = x[n * C + c] - lse;
}
// ClassNLLCriterion:backward (size-averaged) followed by LogSoftMax:updateGradInput
__global__ void k_nll_logsoftmax_bwd(const float* lp, const float* T, float* g, int N, int C) {
  int n = blockIdx.x * blockDim.x + threadIdx.x;
  if (n >= N) return;
  int t = (int)T[n] - 1;
  float sum = -1.f / N;                                  // sum_c dlogp[n][c]
  for (int c = 0; c < C; ++c) g[n * C + c] = (c == t ? -1.f / N : 0.f) - expf(lp[n * C + c]) * sum;
}
__global__ void k_nll_fwd(const float* lp, const float* T, float* out, int N, int C) {
  if (blockIdx.x || threadIdx.x) return;
  float a = 0.f;
  for (int n = 0; n < N; ++n) a -= lp[n * C + (int)T[n] - 1];
  *out = a / N;
}
__global__ void k_axpy(float* x, const float* g, float a, int n) {        // optim.sgd: x += -lr * g
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) x[i] += a * g[i];
}
// keeps stream 0 busy for a while before a foreign producer kernel, so that an unordered consumer would
// actually run ahead of its input (only meaningful in the "nonblocking" demonstration mode)
__global__ void k_delay(long long cycles) {
  long long t0 = clock64();
  while (clock64() - t0 < cycles) {}
}

static inline int gr(int n) { return (n + 127) / 128; }

// ------------------------------------------------------------------ fixture ----------------------
struct Arr { std::vector<uint32_t> dims; std::vector<float> v; };
static std::map<std::string, Arr> load_fixture(const char* path) {
  FILE* f = fopen(path, "rb");
  if (!f) { fprintf(stderr, "cannot open %s\n", path); exit(2); }
  char magic[4]; uint32_t count = 0;
  if (fread(magic, 1, 4, f) != 4 || memcmp(magic, "VBRP", 4) != 0 || fread(&count, 4, 1, f) != 1) { fprintf(stderr, "bad fixture\n"); exit(2); }
  std::map<std::string, Arr> m;
  for (uint32_t i = 0; i < count; ++i) {
    uint32_t nl = 0, nd = 0;
    if (fread(&nl, 4, 1, f) != 1) exit(2);
    std::string name(nl, '\0');
    if (fread(&name[0], 1, nl, f) != nl || fread(&nd, 4, 1, f) != 1) exit(2);
    Arr a; a.dims.resize(nd);
    size_t n = 1;
    for (uint32_t d = 0; d < nd; ++d) { if (fread(&a.dims[d], 4, 1, f) != 1) exit(2); n *= a.dims[d]; }
    a.v.resize(n);
    if (fread(a.v.data(), 4, n, f) != n) exit(2);
    m[name] = a;
  }
  fclose(f);
  return m;
}
static float* upload(const std::vector<float>& v) {
  float* d; CK(cudaMalloc(&d, v.size() * 4));
  CK(cudaMemcpy(d, v.data(), v.size() * 4, cudaMemcpyHostToDevice));      // inputs:cuda() (main.lua:23-24): synchronous
  return d;
}
static double rel_err(const float* dev, const std::vector<float>& want) {
  std::vector<float> h(want.size());
  CK(cudaMemcpy(h.data(), dev, want.size() * 4, cudaMemcpyDeviceToHost));
  double num = 0, den = 0;
  for (size_t i = 0; i < want.size(); ++i) { double d = (double)h[i] - want[i]; num += d * d; den += (double)want[i] * want[i]; }
  return sqrt(num / (den > 1e-60 ? den : 1e-60));
}

int main(int argc, char** argv) {
  if (argc < 2) { fprintf(stderr, "usage: %s <fixture.bin> [legacy|blocking|nonblocking]\n", argv[0]); return 2; }
  const char* mode_s = argc > 2 ? argv[2] : "legacy";
  int mode = VBNN_CTX_STREAM_LEGACY_DEFAULT;
  if (!strcmp(mode_s, "blocking")) mode = VBNN_CTX_STREAM_PRIVATE_BLOCKING;
  else if (!strcmp(mode_s, "nonblocking")) mode = VBNN_CTX_STREAM_GIVEN;
  else if (strcmp(mode_s, "legacy")) { fprintf(stderr, "unknown stream mode %s\n", mode_s); return 2; }
  auto fx = load_fixture(argv[1]);
  auto sc = [&](const char* k) { return fx.at(k).v[0]; };
  const int I0 = (int)fx.at("sizes").v[0], H1 = (int)fx.at("sizes").v[1], H2 = (int)fx.at("sizes").v[2], C = (int)fx.at("sizes").v[3];
  const int S = (int)sc("S"), N = (int)sc("N"), steps = (int)sc("steps");
  CK(cudaSetDevice(0));

  vbnn_ctx* ctx = nullptr;
  VB(vbnn_ctx_create_ex(0, nullptr, mode, 3, &ctx));                     // require 'cunn' (VBLinear.lua:2)
  vbnn_opts o;
  vbnn_opts_default(&o);                                                // config.lua as shipped ...
  o.S = S; o.B = sc("B"); o.mu_init = 1.f; o.var_init = 0.01f;          // ... with the fixture's overrides
  vbnn_layer* vb[2];
  VB(vbnn_layer_create(ctx, I0, H1, VBNN_KIND_VB, &o, &vb[0]));          // mlp.lua:14
  VB(vbnn_layer_create(ctx, H1, H2, VBNN_KIND_VB, &o, &vb[1]));          // mlp.lua:22
  const int dI[3] = {I0, H1, H2}, dO[3] = {H1, H2, C};

  // ---- getParameters() (mlp.lua:37): one flat storage for {weight, bias} of every module, one for the grads
  size_t off_w[3], off_b[3], total = 0;
  for (int k = 0; k < 3; ++k) { off_w[k] = total; total += (size_t)dO[k] * dI[k]; off_b[k] = total; total += dO[k]; }
  float *params, *grads;
  CK(cudaMalloc(&params, total * 4)); CK(cudaMalloc(&grads, total * 4));
  CK(cudaMemset(params, 0, total * 4)); CK(cudaMemset(grads, 0, total * 4));
  for (int k = 0; k < 2; ++k) {
    VB(vbnn_layer_bind(vb[k], VBNN_BUF_WEIGHT, params + off_w[k]));
    VB(vbnn_layer_bind(vb[k], VBNN_BUF_BIAS, params + off_b[k]));
    VB(vbnn_layer_bind(vb[k], VBNN_BUF_GRAD_WEIGHT, grads + off_w[k]));
    VB(vbnn_layer_bind(vb[k], VBNN_BUF_GRAD_BIAS, grads + off_b[k]));
  }
  // the views the module now reports ARE the flat storage
  for (int k = 0; k < 2; ++k) {
    float* p = nullptr; size_t n = 0;
    VB(vbnn_layer_device_ptr(vb[k], VBNN_BUF_GRAD_WEIGHT, &p, &n));
    if (p != grads + off_w[k] || n != (size_t)dO[k] * dI[k]) { fprintf(stderr, "bind: gradWeight view does not follow the flat storage\n"); return 1; }
  }
  // initial parameters of the fixture (mlp.lua:47-55 re-initialises; here the fixture's values)
  char key[64];
  for (int k = 0; k < 2; ++k) {
    snprintf(key, sizeof key, "means0_%d", k); VB(vbnn_layer_set(vb[k], VBNN_BUF_MEANS, fx.at(key).v.data()));
    snprintf(key, sizeof key, "lvars0_%d", k); VB(vbnn_layer_set(vb[k], VBNN_BUF_LVARS, fx.at(key).v.data()));
    float mu_hat, var_hat;
    VB(vbnn_layer_compute_prior(vb[k], &mu_hat, &var_hat));              // VBLinear.lua:46
  }
  CK(cudaMemcpy(params + off_w[2], fx.at("wout0").v.data(), (size_t)C * H2 * 4, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(params + off_b[2], fx.at("bout0").v.data(), (size_t)C * 4, cudaMemcpyHostToDevice));

  // activations owned by the host program (module.output / module.gradInput tensors in Torch)
  float *h[3], *a[2], *logp, *glog, *gh[2], *ga[2], *gx, *derr;
  CK(cudaMalloc(&h[0], N * H1 * 4)); CK(cudaMalloc(&a[0], N * H1 * 4));
  CK(cudaMalloc(&h[1], N * H2 * 4)); CK(cudaMalloc(&a[1], N * H2 * 4));
  CK(cudaMalloc(&h[2], N * C * 4)); CK(cudaMalloc(&logp, N * C * 4)); CK(cudaMalloc(&glog, N * C * 4));
  CK(cudaMalloc(&ga[1], N * H2 * 4)); CK(cudaMalloc(&gh[1], N * H2 * 4));
  CK(cudaMalloc(&ga[0], N * H1 * 4)); CK(cudaMalloc(&gh[0], N * H1 * 4));
  CK(cudaMalloc(&gx, N * I0 * 4)); CK(cudaMalloc(&derr, 4));
  const long long delay = mode == VBNN_CTX_STREAM_GIVEN ? 2000000 : 200000;   // ~0.1 ms of stream-0 work before producers

  bool ok = true;
  for (int it = 0; it < steps; ++it) {
    snprintf(key, sizeof key, "X_%d", it); float* X = upload(fx.at(key).v);
    snprintf(key, sizeof key, "T_%d", it); const std::vector<float>& Th = fx.at(key).v; float* T = upload(Th);
    // ---- net:resetGradients() (mlp.lua:62-67) ----
    CK(cudaMemsetAsync(grads, 0, total * 4, 0));                          // gradParameters:zero()  [foreign]
    for (int k = 0; k < 2; ++k) VB(vbnn_layer_reset_acc(vb[k]));           // VBLinear:resetAcc
    double serr = 0, sacc = 0;
    for (int s = 0; s < S; ++s) {
      // ---- net:sample() (mlp.lua:69-74) with the fixture's epsilon ----
      float* eps[2];
      for (int k = 0; k < 2; ++k) {
        snprintf(key, sizeof key, "eps_%d_%d_%d", it, s, k);
        eps[k] = upload(fx.at(key).v);                                    // e:cuda() (VBLinear.lua:57)
        VB(vbnn_layer_sample(vb[k], s, eps[k]));
      }
      // ---- model:forward (mlp.lua:77) ----
      VB(vbnn_layer_forward(vb[0], X, N, h[0], nullptr));
      k_delay<<<1, 1>>>(delay);
      k_relu<<<gr(N * H1), 128>>>(h[0], a[0], N * H1);                    // [foreign, stream 0]
      VB(vbnn_layer_forward(vb[1], a[0], N, h[1], nullptr));             // must see the ReLU output
      k_relu<<<gr(N * H2), 128>>>(h[1], a[1], N * H2);
      k_linear_fwd<<<gr(N * C), 128>>>(a[1], params + off_w[2], params + off_b[2], h[2], N, H2, C);
      k_logsoftmax<<<gr(N), 128>>>(h[2], logp, N, C);
      // ---- criterion:backward + model:backward (mlp.lua:78-79) ----
      k_nll_logsoftmax_bwd<<<gr(N), 128>>>(logp, T, glog, N, C);
      k_linear_bwd_data<<<gr(N * H2), 128>>>(glog, params + off_w[2], ga[1], N, H2, C);
      k_linear_acc<<<gr(C * H2), 128>>>(a[1], glog, grads + off_w[2], grads + off_b[2], N, H2, C);
      k_delay<<<1, 1>>>(delay);
      k_relu_bwd<<<gr(N * H2), 128>>>(a[1], ga[1], gh[1], N * H2);
      VB(vbnn_layer_backward_data(vb[1], a[0], gh[1], N, ga[0]));        // must see the ReLU backward output
      VB(vbnn_layer_acc_grad(vb[1], a[0], gh[1], N, 1.f));               // VBLinear.lua:112-118
      k_relu_bwd<<<gr(N * H1), 128>>>(a[0], ga[0], gh[0], N * H1);       // must see the library's gradInput
      VB(vbnn_layer_backward_data(vb[0], X, gh[0], N, gx));              // the reference computes it (unused)
      VB(vbnn_layer_acc_grad(vb[0], X, gh[0], N, 1.f));
      // ---- criterion:forward + get_accuracy (mlp.lua:80-82; utils.lua:11-27 host loop) ----
      k_nll_fwd<<<1, 1>>>(logp, T, derr, N, C);
      float err; CK(cudaMemcpy(&err, derr, 4, cudaMemcpyDeviceToHost));
      std::vector<float> lp((size_t)N * C);
      CK(cudaMemcpy(lp.data(), logp, lp.size() * 4, cudaMemcpyDeviceToHost));
      int correct = 0;
      for (int n = 0; n < N; ++n) {
        int arg = 0;
        for (int c = 1; c < C; ++c) if (lp[n * C + c] > lp[n * C + arg]) arg = c;
        correct += (arg + 1 == (int)Th[n]);
      }
      serr += err; sacc += 100.0 * correct / N;
      for (int k = 0; k < 2; ++k) { CK(cudaDeviceSynchronize()); CK(cudaFree(eps[k])); }
    }
    // ---- net:update(opt) (mlp.lua:117-142) ----
    const int n_out = C * H2 + C;
    k_delay<<<1, 1>>>(delay);
    k_axpy<<<gr(n_out), 128>>>(params + off_w[2], grads + off_w[2], -o.lr_bias, n_out);   // optim.sgd on (p, g)
    for (int k = 0; k < 2; ++k) VB(vbnn_layer_update(vb[k], nullptr));                       // VBLinear.lua:124-166
    snprintf(key, sizeof key, "err_%d", it); const double e_want = sc(key);
    snprintf(key, sizeof key, "acc_%d", it); const double a_want = sc(key);
    const double e_got = serr / S, a_got = sacc / S;
    printf("minibatch %d: err %.6f (fixture %.6f)  acc %.3f (fixture %.3f)\n", it, e_got, e_want, a_got, a_want);
    ok &= fabs(e_got - e_want) < 1e-4 * fabs(e_want) && fabs(a_got - a_want) < 1e-3;
    CK(cudaDeviceSynchronize());
    CK(cudaFree(X)); CK(cudaFree(T));
  }
  // ---- final state vs the fixture: library-owned means / lvars, and the BOUND bias in the flat storage ----
  for (int k = 0; k < 2; ++k) {
    float* p; size_t n;
    snprintf(key, sizeof key, "means1_%d", k);
    VB(vbnn_layer_device_ptr(vb[k], VBNN_BUF_MEANS, &p, &n)); double e1 = rel_err(p, fx.at(key).v);
    snprintf(key, sizeof key, "lvars1_%d", k);
    VB(vbnn_layer_device_ptr(vb[k], VBNN_BUF_LVARS, &p, &n)); double e2 = rel_err(p, fx.at(key).v);
    snprintf(key, sizeof key, "bias1_%d", k);
    double e3 = rel_err(params + off_b[k], fx.at(key).v);
    printf("layer %d: means %.2e  lvars %.2e  bias(flat storage) %.2e\n", k, e1, e2, e3);
    ok &= e1 < 2e-4 && e2 < 2e-4 && e3 < 1e-4;
  }
  double ew = rel_err(params + off_w[2], fx.at("wout1").v), eb = rel_err(params + off_b[2], fx.at("bout1").v);
  printf("output nn.Linear: weight %.2e  bias %.2e\n", ew, eb);
  ok &= ew < 1e-5 && eb < 1e-4;
  // hand the buffers back before the flat storage goes away
  for (int k = 0; k < 2; ++k) {
    VB(vbnn_layer_bind(vb[k], VBNN_BUF_WEIGHT, nullptr)); VB(vbnn_layer_bind(vb[k], VBNN_BUF_BIAS, nullptr));
    VB(vbnn_layer_bind(vb[k], VBNN_BUF_GRAD_WEIGHT, nullptr)); VB(vbnn_layer_bind(vb[k], VBNN_BUF_GRAD_BIAS, nullptr));
    VB(vbnn_layer_destroy(vb[k]));
  }
  VB(vbnn_ctx_destroy(ctx));
  printf("REPLAY %s (stream mode: %s)\n", ok ? "OK" : "FAILED", mode_s);
  return ok ? 0 : 1;
}
