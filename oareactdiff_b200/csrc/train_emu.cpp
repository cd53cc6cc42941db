// Host-emulation build of the differentiable training core (train_core.h, -DOARD_HOST_EMU): test infrastructure for
// tests/test_train_emu.py.  Plain C ABI over host pointers; never loaded by the product path.
#define OARD_HOST_EMU
#include "train_core.h"

using namespace oard_train;

static Ctx g_ctx;

extern "C" void emu_reset(int N, int E, int H, int R, int C, int L, int reflect) {
  g_ctx.release();
  g_ctx.W.clear(); g_ctx.wn.clear();
  g_ctx.N = N; g_ctx.E = E; g_ctx.H = H; g_ctx.R = R; g_ctx.C = C; g_ctx.L = L; g_ctx.reflect = reflect;
}
// weights are referenced, not copied: the caller keeps them alive
extern "C" void emu_set_weight(const char* name, float* data, long numel) {
  g_ctx.W[name] = data;
  g_ctx.wn[name] = (size_t)numel;
  if (g_ctx.dW.count(name)) dev_free(g_ctx.dW[name]);
  g_ctx.dW[name] = dev_alloc((size_t)numel);
}
extern "C" void emu_forward_backward(const int* ei, const int* ej, const float* frame, const float* rb, const float* rbf,
                                     const float* inv_deg_i, const float* nodeframe, const float* pos_prjt,
                                     const int* act_idx, int n_act, const float* h_in, float* h_out, float* dpos, const float* g_hout, const float* g_dpos,
                                     float* g_hin) {
  Geometry G{ei, ej, frame, rb, rbf, inv_deg_i, nodeframe, pos_prjt, act_idx, n_act};
  forward(g_ctx, G, h_in, h_out, dpos);
  if (g_hout) {
    for (auto& kv : g_ctx.dW) dev_zero(nullptr, kv.second, g_ctx.wn[kv.first]);
    backward(g_ctx, G, h_in, g_hout, g_dpos, g_hin);
  }
}
extern "C" int emu_get_grad(const char* name, float* dst, long numel) {
  auto it = g_ctx.dW.find(name);
  if (it == g_ctx.dW.end() || (size_t)numel != g_ctx.wn[name]) return -1;
  memcpy(dst, it->second, (size_t)numel * sizeof(float));
  return 0;
}
extern "C" int emu_get_act(const char* name, float* dst, long numel) {
  auto it = g_ctx.act.find(name);
  if (it == g_ctx.act.end() || (size_t)numel > g_ctx.actn[name]) return -1;
  memcpy(dst, it->second, (size_t)numel * sizeof(float));
  return 0;
}
