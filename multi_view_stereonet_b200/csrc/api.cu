// C ABI (include/b200mvs.h): handle, weight packing, workspace and the forward orchestration.
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <mutex>
#include <string>
#include <vector>

#include "../../include/b200mvs.h"
#include "conv.cuh"
#include "conv_tc.cuh"
#include "conv_ws.cuh"
#include "conv5_tc.cuh"
#include "cvf_tc.cuh"
#include "evalpost.cuh"
#include "kernels.cuh"
#include "recurrence.cuh"
#include "tail.cuh"

namespace b200mvs {

static thread_local std::string g_error;
static thread_local int64_t g_launches = 0;
void set_error(const std::string& msg) { g_error = msg; }
void note_launch() { ++g_launches; }
static bool g_pdl = true;
bool pdl_enabled() { return g_pdl; }
void set_pdl_enabled(bool on) { g_pdl = on; }

namespace {
std::mutex g_attr_mutex;
std::map<std::pair<const void*, int>, size_t> g_func_smem;       // (kernel, device) -> opted-in dynamic smem bytes
std::map<std::pair<const void*, int>, bool> g_func_cluster;      // (kernel, device) -> non-portable clusters allowed
std::map<int, int> g_sm_count;                                   // device -> SMs
std::map<std::pair<int, int>, int> g_cluster_size;               // (device, tiles) -> schedulable cluster size
}  // namespace

int ensure_func_smem(const void* func, size_t bytes) {
  int dev = 0;
  B200MVS_CUDA_OK(cudaGetDevice(&dev));
  std::lock_guard<std::mutex> lock(g_attr_mutex);
  size_t& have = g_func_smem[{func, dev}];
  if (bytes > have) {
    B200MVS_CUDA_OK(cudaFuncSetAttribute(func, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes));
    have = bytes;
  }
  return 0;
}
int ensure_func_nonportable_cluster(const void* func) {
  int dev = 0;
  B200MVS_CUDA_OK(cudaGetDevice(&dev));
  std::lock_guard<std::mutex> lock(g_attr_mutex);
  bool& have = g_func_cluster[{func, dev}];
  if (!have) {
    B200MVS_CUDA_OK(cudaFuncSetAttribute(func, cudaFuncAttributeNonPortableClusterSizeAllowed, 1));
    have = true;
  }
  return 0;
}
// SMs the persistent one-CTA-per-SM kernels (conv_ws, cvf_tc) size their grids for.  While another lane's depth sweep
// holds clusters, a grid of all SMs cannot be resident at once and its statically assigned tiles would run as two
// waves; in lane mode the grids are sized for the SMs the sweeps leave free (b200mvs_forward).
static thread_local int g_sm_reserved = 0;
void set_reserved_sms(int n) { g_sm_reserved = n; }
int current_device_sm_count(int* sms) {
  int dev = 0;
  B200MVS_CUDA_OK(cudaGetDevice(&dev));
  std::lock_guard<std::mutex> lock(g_attr_mutex);
  int& n = g_sm_count[dev];
  if (n == 0) B200MVS_CUDA_OK(cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev));
  *sms = n - g_sm_reserved > 16 ? n - g_sm_reserved : 16;
  return 0;
}
int cached_cluster_size(int tiles) {
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return 0;
  std::lock_guard<std::mutex> lock(g_attr_mutex);
  auto it = g_cluster_size.find({dev, tiles});
  return it == g_cluster_size.end() ? 0 : it->second;
}
void remember_cluster_size(int tiles, int cluster) {
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return;
  std::lock_guard<std::mutex> lock(g_attr_mutex);
  g_cluster_size[{dev, tiles}] = cluster;
}

// Event-pair probe around launches of one kernel class (b200mvs_probe_select).
struct Probe {
  int tag = 0;
  std::vector<cudaEvent_t> ev;  // pairs
  size_t used = 0;
  cudaEvent_t next() {
    if (used == ev.size()) {
      cudaEvent_t e;
      cudaEventCreate(&e);
      ev.push_back(e);
    }
    return ev[used++];
  }
};
static thread_local Probe* g_probe = nullptr;
void probe_before(int tag, cudaStream_t stream) {
  if (g_probe != nullptr && g_probe->tag == tag) cudaEventRecord(g_probe->next(), stream);
}
void probe_after(int tag, cudaStream_t stream) {
  if (g_probe != nullptr && g_probe->tag == tag) cudaEventRecord(g_probe->next(), stream);
}

namespace {

// Debug (B200MVS_LANE_TRACE=1): device timeline of a multi-lane forward, one line per lane with the time of every
// stage boundary relative to the start of the call.
struct TraceMark { int lane; const char* what; cudaEvent_t ev; };
std::vector<TraceMark> g_trace;
cudaEvent_t g_trace_origin = nullptr;
bool lane_trace_enabled() {
  static const bool on = getenv("B200MVS_LANE_TRACE") != nullptr;
  return on;
}

struct ConvW {
  float* w = nullptr;
  float* bias = nullptr;
  uint8_t* w16 = nullptr;   // fp16 UMMA blocks (3x3 layers that may run on tensor cores)
  uint8_t* w16s = nullptr;  // split hi/lo fp16 UMMA blocks (fp32-class accuracy)
};
struct GnW {
  float* gamma = nullptr;
  float* beta = nullptr;
};

struct Refiner {
  ConvW conv0, res[6], fin;
  GnW gn0, gn[6];
  RefineFinalW finw;
  RefineHeadIdW idw;   // conv0's idepth-channel taps (the last input channel), for the precomputed-guide path
};


// Bump allocator over one device allocation; sizes are computed from the call shape.
struct Arena {
  char* base = nullptr;
  size_t capacity = 0, used = 0;
  bool dry = false;
  template <class T>
  T* take(size_t count) {
    const size_t bytes = (count * sizeof(T) + 255) & ~(size_t)255;
    T* p = dry ? nullptr : reinterpret_cast<T*>(base + used);
    used += bytes;
    return p;
  }
};

struct Workspace {
  // geometry
  GeomOut geo{};
  // level-0 warp
  float* warped0 = nullptr;
  uint8_t* l0mask = nullptr;
  // feature network ((1+V)*B images; left images first)
  float *f1 = nullptr, *f2 = nullptr, *f3 = nullptr, *feat4 = nullptr;
  float *l4x[2] = {nullptr, nullptr}, *l4y[2] = {nullptr, nullptr};
  // recurrence
  float* imgconv = nullptr;
  float* rec_plan = nullptr;   // gather plan of the persistent sweep, [n][D][recurrence_plan_stride][4]
  int* rec_flags = nullptr;    // [n][17] progress / out-of-window flags of the persistent sweep
  WideScratch wide;   // scratch of the wide sweep (sweep_wide.cu)
  float *vol = nullptr, *wf = nullptr, *wimg = nullptr, *sy0 = nullptr, *sx0 = nullptr, *sy1 = nullptr;
  // cost volume
  float *cost = nullptr, *cvfA = nullptr, *cost1 = nullptr;
  uint8_t* mask_views = nullptr;
  float *raw_views = nullptr, *refined_views = nullptr;
  // refiners (shared by all levels)
  float *rx[2] = {nullptr, nullptr}, *ry[2] = {nullptr, nullptr};
  // guide part of every refiner's conv0 (+ bias), computed while the depth sweep runs: [B][H_l][W_l][32] fp32
  float* pre[5] = {nullptr, nullptr, nullptr, nullptr, nullptr};
  // output scratch
  float* idepth[5] = {nullptr, nullptr, nullptr, nullptr, nullptr};
  float* prior[5] = {nullptr, nullptr, nullptr, nullptr, nullptr};
  uint8_t* mask[5] = {nullptr, nullptr, nullptr, nullptr, nullptr};
  // GroupNorm statistics arena
  double* stats = nullptr;
  size_t stats_count = 0;
};

}  // namespace
}  // namespace b200mvs

using namespace b200mvs;

// Everything one in-flight forward owns: workspace, side stream, fork / join events.  A call whose pairs need more than
// one round of the depth sweep's clusters is cut into lanes of whole image groups that run concurrently on their own
// streams (b200mvs_forward): while the sweep of one lane holds its 11-SM clusters, the other lanes' kernels fill the
// rest of the chip.
constexpr int kRecProfWords = 16 * 12 + 16 * 16 * 32;   // recurrence phase totals + one step's per-warp timeline
constexpr int kMaxLanes = 2;   // 3 streams per lane + the caller's: more would exceed the 8 hardware queues of a device
                               // (CUDA_DEVICE_MAX_CONNECTIONS) and serialise lanes behind each other (measured)
struct Lane {
  Arena arena;
  Workspace ws;
  b200mvs_shape ws_shape{};
  bool ws_valid = false;
  cudaStream_t main = nullptr;   // lanes >= 1 run here (lane 0 runs on the caller's stream)
  cudaStream_t side = nullptr;
  cudaStream_t rec = nullptr;    // high priority: the depth sweep's clusters are placed before anything else pending
  cudaEvent_t ev_pre = nullptr, ev_right = nullptr, ev_geo = nullptr, ev_imgconv = nullptr, ev_fork = nullptr,
              ev_left = nullptr, ev_mask_in = nullptr, ev_mask_out = nullptr, ev_rec_in = nullptr, ev_rec_out = nullptr,
              ev_end = nullptr;
};

struct b200mvs_net {
  int device = 0;
  std::vector<void*> allocs;  // weight allocations
  // FeatureNetwork (multi_view_stereonet.py:78-129)
  ConvW feat_conv[4];
  ConvW feat_res[6];
  GnW feat_gn[6];
  ConvW feat_final;
  // FeatureRefiner (multi_view_stereonet.py:398-440)
  ConvW fr_conv0, fr_res0, fr_final;
  GnW fr_gn0, fr_gn1;
  uint8_t* fr_w16 = nullptr;  // split-fp16 weights of the three layers for the persistent kernel
  // CostVolumeFilter (multi_view_stereonet.py:302-353)
  ConvW cvf[5];
  GnW cvf_gn[4];
  CvfFinalW cvf_finw;
  RefineHeadW head0;   // refiner0.conv0 (level 0: image + idepth only)
  // IDepthmapRefiner x5 (multi_view_stereonet.py:442-484)
  Refiner refiner[5];

  Lane lanes[kMaxLanes];
  Lane* cur = &lanes[0];          // the lane whose kernels are being enqueued (host enqueue is sequential)
  int lanes_max = 0;              // option "lanes": 0 = automatic (plan_lanes), 1 = never split a call, 2 = up to two
                                  // concurrent lanes whenever the sweep needs more than one round of clusters
  int last_lanes = 1;
  cudaEvent_t ev_begin = nullptr;
  Probe probe;
  // device staging of b200mvs_forward_host (inputs, outputs), grown on demand
  char* host_stage = nullptr;
  size_t host_stage_bytes = 0;
  bool keep_stages = false;
  // Operand precision of the refiner convolutions at the large levels: fp16 operands cost ~0.05-0.15 px of
  // disparity-equivalent error on idepth * fx, which is 1e-4 of an idepth range of 63 px (64 hypotheses) but 1-3e-3 of
  // a range of 11 px (12 hypotheses, the reference's own training configuration) -- beyond the 1e-3 parity bar.
  // 2 (default) = split fp16 at every level when the sweep has fewer than 48 hypotheses, 1 = always, 0 = never.
  int precise_refiners = 2;
  bool precise_now = false;
  bool use_tensor_cores = true;
  bool half_activations = true;
  bool warp_specialized = true;
  bool early_d2h = true;          // forward_host: levels 1-4 downloaded on the copy stream next to the level-0 refiner
  bool left_late = true;          // left feature network waits for the right one (side stream)
  bool conv0_precompute = true;   // refiner conv0 = precomputed guide part + idepth part (tail.cu)
  int rec_debug = 0;
  unsigned* wide_abort_host = nullptr;   // pinned: the wide sweep's watchdog flag of the last forward that used it
  int sweep_mode = 0;             // option "sweep": 0 = cluster kernel where the image fits one cluster, else the wide
                                  // kernel; 1 = wide kernel wherever it is supported; 2 = step by step (a launch per layer)
  // Side stream for the work that does not depend on the comparison views (left feature network) or that
  // nothing downstream waits for (mask upsampling): forked / joined with events inside one forward.
  cudaEvent_t ev_coarse = nullptr;
  bool overlap = true;
  // b200mvs_forward_host: uploads run on their own stream in the order the path needs them, compute waits per piece
  cudaStream_t copy_stream = nullptr, host_stream = nullptr;
  cudaEvent_t ev_up[5] = {nullptr, nullptr, nullptr, nullptr, nullptr};   // cameras, right l0, right l4, left l0, rest
  float* pinned_small = nullptr;
  size_t pinned_small_floats = 0;
  long long* rec_prof = nullptr;  // device [16][12] phase totals + [16][16][32] per-warp timeline of one step, allocated
                                  // when option "recurrence_profile" is set
  bool l4_chain = true;             // option "l4_chain": the level-4 tail of the feature network as one cluster kernel
  bool prio_main = false;           // option "prio_main": single-lane forwards run on a high-priority stream
  bool stage_profile = false;       // option "stage_profile": events at the stage boundaries of the main stream
  std::string last_stage_profile;   // "name=us;..." of the last profiled forward (b200mvs_last_stage_profile)
  b200mvs_shape last_shape{};
  bool have_last = false;
  int64_t last_launches = 0;
};

namespace {

struct StateDict {
  std::map<std::string, std::pair<const float*, int64_t>> t;
  const float* get(const std::string& name, int64_t numel) const {
    auto it = t.find(name);
    if (it == t.end()) {
      set_error("state dict has no tensor '" + name + "'");
      return nullptr;
    }
    if (it->second.second != numel) {
      set_error("tensor '" + name + "' has " + std::to_string(it->second.second) + " elements, expected " +
                std::to_string(numel));
      return nullptr;
    }
    return it->second.first;
  }
};

int upload(b200mvs_net* net, const std::vector<float>& host, float** dev) {
  B200MVS_CUDA_OK(cudaMalloc(dev, host.size() * sizeof(float)));
  net->allocs.push_back(*dev);
  B200MVS_CUDA_OK(cudaMemcpy(*dev, host.data(), host.size() * sizeof(float), cudaMemcpyHostToDevice));
  return 0;
}

// Packs a reference conv weight (O, I, taps) into [chunk][tap][8][O]: four chunks for the 32
// feature channels (reference input index feat_off + c) if has_feat, then one zero-padded chunk for
// the planar "extra" channels (reference indices extra_idx[0..n_extra)).
int pack_conv(b200mvs_net* net, const StateDict& sd, const std::string& name, int cout, int cin, int taps,
              bool has_feat, int feat_off, const std::vector<int>& extra_idx, bool has_bias, ConvW* out) {
  const float* w = sd.get(name + ".weight", (int64_t)cout * cin * taps);
  if (w == nullptr) return B200MVS_EWEIGHTS;
  const int chunks = (has_feat ? 4 : 0) + (extra_idx.empty() ? 0 : 1);
  std::vector<float> packed((size_t)chunks * taps * 8 * cout, 0.f);
  for (int chunk = 0; chunk < chunks; ++chunk)
    for (int tap = 0; tap < taps; ++tap)
      for (int k = 0; k < 8; ++k) {
        int ci;
        if (has_feat && chunk < 4) {
          ci = feat_off + chunk * 8 + k;
        } else {
          if (k >= (int)extra_idx.size()) continue;
          ci = extra_idx[k];
        }
        for (int o = 0; o < cout; ++o)
          packed[(((size_t)chunk * taps + tap) * 8 + k) * cout + o] = w[((size_t)o * cin + ci) * taps + tap];
      }
  int rc = upload(net, packed, &out->w);
  if (rc) return rc;
  if (has_bias) {
    const float* b = sd.get(name + ".bias", cout);
    if (b == nullptr) return B200MVS_EWEIGHTS;
    std::vector<float> hb(b, b + cout);
    rc = upload(net, hb, &out->bias);
    if (rc) return rc;
  }
  return 0;
}

int pack_conv_tc(b200mvs_net* net, const StateDict& sd, const std::string& name, int cin, bool has_feat,
                 int feat_off, const std::vector<int>& extra_idx, ConvW* out) {
  const float* w = sd.get(name + ".weight", (int64_t)32 * cin * 9);
  if (w == nullptr) return B200MVS_EWEIGHTS;
  for (int split = 0; split < 2; ++split) {
    std::vector<uint8_t> packed;
    pack_conv3x3_tc_weights(w, cin, has_feat, feat_off, extra_idx, split != 0, &packed);
    uint8_t** dst = split ? &out->w16s : &out->w16;
    B200MVS_CUDA_OK(cudaMalloc(reinterpret_cast<void**>(dst), packed.size()));
    net->allocs.push_back(*dst);
    B200MVS_CUDA_OK(cudaMemcpy(*dst, packed.data(), packed.size(), cudaMemcpyHostToDevice));
  }
  return 0;
}

int pack_gn(b200mvs_net* net, const StateDict& sd, const std::string& name, GnW* out) {
  const float* g = sd.get(name + ".weight", kC);
  const float* b = sd.get(name + ".bias", kC);
  if (g == nullptr || b == nullptr) return B200MVS_EWEIGHTS;
  int rc = upload(net, std::vector<float>(g, g + kC), &out->gamma);
  if (rc) return rc;
  return upload(net, std::vector<float>(b, b + kC), &out->beta);
}

#define RC(expr)            \
  do {                      \
    int _rc = (expr);       \
    if (_rc != 0) return _rc; \
  } while (0)

int build_weights(b200mvs_net* net, const StateDict& sd) {
  const std::string fe = "left_feature_extractor";
  RC(pack_conv(net, sd, fe + ".conv0", 32, 3, 25, false, 0, {0, 1, 2}, false, &net->feat_conv[0]));
  for (int i = 1; i < 4; ++i)
    RC(pack_conv(net, sd, fe + ".conv" + std::to_string(i), 32, 32, 25, true, 0, {}, false, &net->feat_conv[i]));
  for (int i = 0; i < 4; ++i) {
    const int cin = i == 0 ? 3 : 32;
    const float* w = sd.get(fe + ".conv" + std::to_string(i) + ".weight", (int64_t)32 * cin * 25);
    if (w == nullptr) return B200MVS_EWEIGHTS;
    std::vector<uint8_t> packed;
    if (i == 0) pack_conv5_c3_weights(w, &packed);
    else pack_conv5_c32_weights(w, &packed);
    B200MVS_CUDA_OK(cudaMalloc(reinterpret_cast<void**>(&net->feat_conv[i].w16s), packed.size()));
    net->allocs.push_back(net->feat_conv[i].w16s);
    B200MVS_CUDA_OK(cudaMemcpy(net->feat_conv[i].w16s, packed.data(), packed.size(), cudaMemcpyHostToDevice));
  }
  for (int i = 0; i < 6; ++i) {
    const std::string r = fe + ".res" + std::to_string(i);
    RC(pack_conv(net, sd, r + ".conv1", 32, 32, 9, true, 0, {}, false, &net->feat_res[i]));
    RC(pack_conv_tc(net, sd, r + ".conv1", 32, true, 0, {}, &net->feat_res[i]));
    RC(pack_gn(net, sd, r + ".bn1", &net->feat_gn[i]));
  }
  RC(pack_conv(net, sd, fe + ".conv_final", 32, 32, 9, true, 0, {}, true, &net->feat_final));
  RC(pack_conv_tc(net, sd, fe + ".conv_final", 32, true, 0, {}, &net->feat_final));

  const std::string fr = "right_feature_extractor.refiner";
  // input = cat([image(3), features(32)])  (multi_view_stereonet.py:425)
  RC(pack_conv(net, sd, fr + ".conv0", 32, 35, 9, true, 3, {0, 1, 2}, true, &net->fr_conv0));
  RC(pack_conv_tc(net, sd, fr + ".conv0", 35, true, 3, {0, 1, 2}, &net->fr_conv0));
  RC(pack_gn(net, sd, fr + ".bn0", &net->fr_gn0));
  RC(pack_conv(net, sd, fr + ".res0.conv1", 32, 32, 9, true, 0, {}, true, &net->fr_res0));
  RC(pack_conv_tc(net, sd, fr + ".res0.conv1", 32, true, 0, {}, &net->fr_res0));
  RC(pack_gn(net, sd, fr + ".res0.bn1", &net->fr_gn1));
  RC(pack_conv(net, sd, fr + ".conv_final", 32, 32, 9, true, 0, {}, true, &net->fr_final));
  RC(pack_conv_tc(net, sd, fr + ".conv_final", 32, true, 0, {}, &net->fr_final));

  {
    const float* w0 = sd.get(fr + ".conv0.weight", 32 * 35 * 9);
    const float* w1 = sd.get(fr + ".res0.conv1.weight", 32 * 32 * 9);
    const float* w2 = sd.get(fr + ".conv_final.weight", 32 * 32 * 9);
    if (w0 == nullptr || w1 == nullptr || w2 == nullptr) return B200MVS_EWEIGHTS;
    std::vector<uint8_t> packed;
    pack_recurrence_weights(w0, w1, w2, &packed);
    B200MVS_CUDA_OK(cudaMalloc(reinterpret_cast<void**>(&net->fr_w16), packed.size()));
    net->allocs.push_back(net->fr_w16);
    B200MVS_CUDA_OK(cudaMemcpy(net->fr_w16, packed.data(), packed.size(), cudaMemcpyHostToDevice));
  }

  for (int i = 0; i < 4; ++i) {
    {
      const float* w = sd.get("volume_filter4.conv" + std::to_string(i) + ".weight", 32 * 32 * 27);
      if (w == nullptr) return B200MVS_EWEIGHTS;
      std::vector<uint8_t> packed;
      pack_cvf_tc_weights(w, &packed);
      B200MVS_CUDA_OK(cudaMalloc(reinterpret_cast<void**>(&net->cvf[i].w16), packed.size()));
      net->allocs.push_back(net->cvf[i].w16);
      B200MVS_CUDA_OK(cudaMemcpy(net->cvf[i].w16, packed.data(), packed.size(), cudaMemcpyHostToDevice));
    }
    RC(pack_conv(net, sd, "volume_filter4.conv" + std::to_string(i), 32, 32, 27, true, 0, {}, true, &net->cvf[i]));
    RC(pack_gn(net, sd, "volume_filter4.bn" + std::to_string(i), &net->cvf_gn[i]));
  }
  RC(pack_conv(net, sd, "volume_filter4.conv4", 1, 32, 27, true, 0, {}, true, &net->cvf[4]));
  {
    const float* w = sd.get("volume_filter4.conv4.weight", 32 * 27);
    const float* b = sd.get("volume_filter4.conv4.bias", 1);
    if (w == nullptr || b == nullptr) return B200MVS_EWEIGHTS;
    std::memcpy(net->cvf_finw.w, w, sizeof(net->cvf_finw.w));
    net->cvf_finw.bias = b[0];
  }

  for (int lvl = 0; lvl < 5; ++lvl) {
    const std::string r = "refiner" + std::to_string(lvl);
    Refiner& R = net->refiner[lvl];
    // input = cat([image(3), features(32), idepth(1)]) for levels 1..4, cat([image(3), idepth(1)]) at
    // level 0 (multi_view_stereonet.py:469, 610, 679).
    if (lvl > 0) {
      RC(pack_conv(net, sd, r + ".conv0", 32, 36, 9, true, 3, {0, 1, 2, 35}, true, &R.conv0));
      RC(pack_conv_tc(net, sd, r + ".conv0", 36, true, 3, {0, 1, 2, 35}, &R.conv0));
    } else {
      RC(pack_conv(net, sd, r + ".conv0", 32, 4, 9, false, 0, {0, 1, 2, 3}, true, &R.conv0));
      RC(pack_conv_tc(net, sd, r + ".conv0", 4, false, 0, {0, 1, 2, 3}, &R.conv0));
    }
    {
      const int cin = lvl > 0 ? 36 : 4;
      const float* w = sd.get(r + ".conv0.weight", 32 * cin * 9);
      if (w == nullptr) return B200MVS_EWEIGHTS;
      for (int t = 0; t < 9; ++t)
        for (int o = 0; o < 32; ++o) R.idw.w[t * 32 + o] = w[((size_t)o * cin + (cin - 1)) * 9 + t];
    }
    if (lvl == 0) {
      const float* w = sd.get(r + ".conv0.weight", 32 * 4 * 9);
      const float* b = sd.get(r + ".conv0.bias", 32);
      if (w == nullptr || b == nullptr) return B200MVS_EWEIGHTS;
      std::memcpy(net->head0.w, w, sizeof(net->head0.w));
      std::memcpy(net->head0.bias, b, sizeof(net->head0.bias));
    }
    RC(pack_gn(net, sd, r + ".bn0", &R.gn0));
    for (int i = 0; i < 6; ++i) {
      const std::string cn = r + ".res" + std::to_string(i) + ".conv1";
      RC(pack_conv(net, sd, cn, 32, 32, 9, true, 0, {}, true, &R.res[i]));
      RC(pack_conv_tc(net, sd, cn, 32, true, 0, {}, &R.res[i]));
      RC(pack_gn(net, sd, r + ".res" + std::to_string(i) + ".bn1", &R.gn[i]));
    }
    RC(pack_conv(net, sd, r + ".conv_final", 1, 32, 9, true, 0, {}, true, &R.fin));
    {
      const float* w = sd.get(r + ".conv_final.weight", 32 * 9);
      const float* b = sd.get(r + ".conv_final.bias", 1);
      if (w == nullptr || b == nullptr) return B200MVS_EWEIGHTS;
      std::memcpy(R.finw.w, w, sizeof(R.finw.w));
      R.finw.bias = b[0];
    }
  }
  return 0;
}

// The wide sweep's watchdog (sweep_wide.cu: chain_wait): non-zero once a forward's chain barrier has timed out.
int check_sweep_watchdog(b200mvs_net* net) {
  if (net->wide_abort_host == nullptr || *net->wide_abort_host == 0) return 0;
  *net->wide_abort_host = 0;
  set_error("depth sweep: a chain barrier of the wide kernel timed out (its CTAs were not all making progress -- e.g. two "
            "such forwards running concurrently on one GPU); the results of that forward are invalid");
  return B200MVS_ECUDA;
}

struct Levels {
  int h[5], w[5];
  size_t px[5];
};
Levels levels_of(const b200mvs_shape& s) {
  Levels L;
  L.h[0] = s.rows;
  L.w[0] = s.cols;
  for (int l = 1; l < 5; ++l) {
    L.h[l] = (L.h[l - 1] + 1) / 2;
    L.w[l] = (L.w[l - 1] + 1) / 2;
  }
  for (int l = 0; l < 5; ++l) L.px[l] = (size_t)L.h[l] * L.w[l];
  return L;
}

// Lays the workspace out in the arena (dry run computes the size).
void layout(b200mvs_net* net, const b200mvs_shape& s, bool dry) {
  Arena& A = net->cur->arena;
  Workspace& W = net->cur->ws;
  A.used = 0;
  A.dry = dry;
  const Levels L = levels_of(s);
  const size_t B = s.batch, V = s.views, D = s.num_idepth_samples;
  const size_t n = B * V, NI = B + n;
  W.geo.baseline = A.take<float>(n);
  W.geo.samples = A.take<float>(n * D);
  W.geo.H0 = A.take<float>(n * 9);
  W.geo.H = A.take<float>(n * D * 9);
  W.geo.Hinc = A.take<float>(n * D * 9);
  W.warped0 = A.take<float>(n * 3 * L.px[0]);
  W.l0mask = A.take<uint8_t>(n * L.px[0]);
  W.f1 = A.take<float>(NI * L.px[1] * kC);
  W.f2 = A.take<float>(NI * L.px[2] * kC);
  W.f3 = A.take<float>(NI * L.px[3] * kC);
  W.feat4 = A.take<float>(NI * L.px[4] * kC);
  for (int i = 0; i < 2; ++i) {
    W.l4x[i] = A.take<float>(NI * L.px[4] * kC);
    W.l4y[i] = A.take<float>(NI * L.px[4] * kC);
  }
  W.vol = A.take<float>(n * D * L.px[4] * kC);
  W.imgconv = A.take<float>(n * D * L.px[4] * kC);
  if (recurrence_supported(L.h[4], L.w[4], nullptr, nullptr) || sweep_wide_supported(L.h[4], L.w[4]))
    W.rec_plan = A.take<float>(n * D * (size_t)recurrence_plan_stride(L.h[4], L.w[4]) * 4);
  W.rec_flags = A.take<int>(n * 17);
  if (sweep_wide_supported(L.h[4], L.w[4])) {
    size_t wf = 0, y = 0, part = 0, ctr = 0;
    sweep_wide_scratch(L.h[4], L.w[4], (int)n, &wf, &y, &part, &ctr);
    W.wide.wf = A.take<float>(wf);
    W.wide.y = A.take<float>(y);
    W.wide.part = A.take<float>(2 * part);
    W.wide.ctr = A.take<unsigned>(ctr);
  }
  W.wf = A.take<float>(n * L.px[4] * kC);
  W.wimg = A.take<float>(n * 3 * L.px[4]);
  W.sy0 = A.take<float>(n * L.px[4] * kC);
  W.sx0 = A.take<float>(n * L.px[4] * kC);
  W.sy1 = A.take<float>(n * L.px[4] * kC);
  W.cost = net->keep_stages ? A.take<float>(n * D * L.px[4] * kC) : nullptr;
  W.cvfA = A.take<float>(n * D * L.px[4] * kC);
  W.cost1 = A.take<float>(n * D * L.px[4]);
  W.mask_views = A.take<uint8_t>(n * D * L.px[4]);
  W.raw_views = A.take<float>(n * L.px[4]);
  W.refined_views = A.take<float>(n * L.px[4]);
  // refiner ping-pong buffers: B images at level 0 or B*V images at level 4, whichever is larger
  const size_t rmax = (B * L.px[0] > n * L.px[4]) ? B * L.px[0] : n * L.px[4];
  for (int i = 0; i < 2; ++i) {
    W.rx[i] = A.take<float>(rmax * kC);
    W.ry[i] = A.take<float>(rmax * kC);
  }
  for (int l = 0; l < 5; ++l) W.pre[l] = A.take<float>(B * L.px[l] * kC);
  for (int l = 0; l < 5; ++l) {
    W.idepth[l] = A.take<float>(B * L.px[l]);
    W.prior[l] = A.take<float>(B * L.px[l]);
    W.mask[l] = (l == 0) ? nullptr : A.take<uint8_t>(B * D * L.px[l]);
  }
  // statistics: featnet 6 layers x NI, recurrence 2 x (D-1) x n, cvf 4 x n, refiners 7 x (n + 4B)
  W.stats_count = (6 * NI + 2 * (D > 0 ? D - 1 : 0) * n + 4 * n + 7 * (n + 4 * B)) * 2 * kGroups;
  W.stats = A.take<double>(W.stats_count);
}

bool same_shape(const b200mvs_shape& a, const b200mvs_shape& b) {
  return a.batch == b.batch && a.views == b.views && a.rows == b.rows && a.cols == b.cols &&
         a.num_idepth_samples == b.num_idepth_samples;
}

int ensure_workspace(b200mvs_net* net, const b200mvs_shape& s) {
  if (net->cur->ws_valid && same_shape(net->cur->ws_shape, s)) return 0;
  layout(net, s, true);
  const size_t need = net->cur->arena.used;
  if (need > net->cur->arena.capacity) {
    if (net->cur->arena.base != nullptr) {
      B200MVS_CUDA_OK(cudaDeviceSynchronize());
      B200MVS_CUDA_OK(cudaFree(net->cur->arena.base));
      net->cur->arena.base = nullptr;
      net->cur->arena.capacity = 0;
    }
    cudaError_t e = cudaMalloc(reinterpret_cast<void**>(&net->cur->arena.base), need);
    if (e != cudaSuccess) {
      set_error("workspace allocation of " + std::to_string(need) + " bytes failed: " + cudaGetErrorString(e));
      return B200MVS_ENOMEM;
    }
    net->cur->arena.capacity = need;
  }
  layout(net, s, false);
  net->cur->ws_shape = s;
  net->cur->ws_valid = true;
  return 0;
}

// 3x3 convolution with 32 outputs: tensor cores when enabled (fp16 operands on the large levels, split hi/lo
// fp16 on the 1/8- and 1/16-scale levels that need fp32-class accuracy), the fp32 FFMA kernel otherwise.
int conv3x3_c32(b200mvs_net* net, const ConvParams& p, const ConvW& w, bool precise, cudaStream_t stream) {
  if (net->use_tensor_cores && net->warp_specialized && !precise && w.w16 != nullptr && conv3x3_ws_supported(p))
    return launch_conv3x3_ws(p, w.w16, stream);
  if (net->use_tensor_cores && w.w16 != nullptr && conv3x3_tc_supported(p))
    return launch_conv3x3_tc(p, precise ? w.w16s : w.w16, precise, stream);
  return launch_conv(CONV_3x3, 32, p, stream);
}

struct StatsCursor {
  double* base;
  size_t used = 0, cap;
  double* take(size_t imgs) {
    double* p = base + used;
    used += imgs * 2 * kGroups;
    return p;
  }
};

// One IDepthmapRefiner (multi_view_stereonet.py:468-484) with the caller-side fx scaling
// (:607-611) folded into the first loader and the last epilogue.
// `pre` != nullptr: the guide part of conv0 (+ bias) was computed ahead of time by run_refiner_guide (image m reads
// pre[m / guide_div]); conv0 is then only the idepth channel on top of it.
int run_refiner(b200mvs_net* net, const Refiner& R, StatsCursor& sc, int m, int H, int W, const float* guide_feat,
                int guide_div, const float* image, int image_div, const float* prior, const float* Kl, int k_div,
                float* out, int res_tag, cudaStream_t stream, const float* pre = nullptr,
                const float* coarse = nullptr, int ch = 0, int cw = 0) {
  Workspace& ws = net->cur->ws;
  const size_t P = (size_t)H * W;
  const double inv_count = 1.0 / (8.0 * (double)P);
  double* st_prev = nullptr;
  // levels 3 and 4 of a 512x640 input always; every level when the call asks for it (forward_impl: few hypotheses)
  const bool precise = (long long)H * W <= 96 * 128 || net->precise_now;
  // Levels 0-2 on the tensor-core path keep the refiner's internal activations (raw conv outputs and the
  // residual stream) in fp16: halves the HBM traffic of 48 % of the network's bytes; measured cost ~1e-4 of
  // the 1e-3 parity budget.  x_out with half storage is only written by the tensor-core kernel.
  const int half_act = (net->use_tensor_cores && !precise && net->half_activations) ? 1 : 0;

  ConvParams p;
  p.n_img = m;
  p.Hi = p.Ho = H;
  p.Wi = p.Wo = W;
  // conv0
  if (guide_feat != nullptr) {
    p.feat.ptr = guide_feat;
    p.feat.mode = FEAT_RAW;
    p.feat.img_div = guide_div;
  }
  p.extra.n = 4;
  for (int e = 0; e < 3; ++e) {
    p.extra.ptr[e] = image + e * P;
    p.extra.img_stride[e] = 3 * (long long)P;
    p.extra.img_div[e] = image_div;
  }
  p.extra.ptr[3] = prior;
  p.extra.img_stride[3] = (long long)P;
  p.extra.img_div[3] = 1;
  p.extra.scale[3] = Kl;  // fx = K[b][0][0]
  p.extra.scale_div[3] = k_div;
  p.extra.scale_stride[3] = 16;
  p.w = R.conv0.w;
  p.bias = R.conv0.bias;
  p.dil = 1;
  p.out = ws.ry[0];
  p.out_half = half_act;
  p.out_stats = st_prev = sc.take(m);
  // conv0 sees the idepth channel scaled by fx (values of a few hundred with the signal in the low bits):
  // always split precision.
  if (pre != nullptr) {
    // with `coarse` the head also upsamples the coarser level's idepth into `prior` (one launch less per level)
    RC(launch_refine_head_pre(pre, guide_div, prior, coarse, ch, cw, const_cast<float*>(prior), Kl, k_div, 16, R.idw, m, H,
                              W, ws.ry[0], half_act != 0, st_prev, stream));
  } else if (guide_feat == nullptr && image_div == 1 && net->use_tensor_cores) {
    // level 0: four planar inputs only -- dedicated fp32 kernel (tail.cu)
    RC(launch_refine_head_l0(image, prior, Kl, k_div, 16, net->head0, m, H, W, ws.ry[0], half_act != 0, st_prev, stream));
  } else {
    RC(conv3x3_c32(net, p, R.conv0, true, stream));
  }

  static const int dilations[6] = {1, 2, 4, 8, 1, 1};  // multi_view_stereonet.py:457
  const GnW* gn_prev = &R.gn0;
  int ycur = 0;      // ry[ycur] holds the raw output of the previous conv
  int xres = -1;     // rx[xres] holds the previous block's input (the residual), -1 = none
  for (int i = 0; i < 6; ++i) {
    ConvParams q;
    q.n_img = m;
    q.Hi = q.Ho = H;
    q.Wi = q.Wo = W;
    q.feat.ptr = ws.ry[ycur];
    q.feat.mode = (xres < 0) ? FEAT_GN : FEAT_GN_RES;
    q.feat.stats = st_prev;
    q.feat.gamma = gn_prev->gamma;
    q.feat.beta = gn_prev->beta;
    q.feat.inv_count = inv_count;
    q.feat.resid = (xres < 0) ? nullptr : ws.rx[xres];
    const int xnew = (xres < 0) ? 0 : 1 - xres;
    q.feat.x_out = ws.rx[xnew];
    q.feat.half_io = half_act;
    q.out_half = half_act;
    q.w = R.res[i].w;
    q.bias = R.res[i].bias;
    q.dil = dilations[i];
    q.out = ws.ry[1 - ycur];
    q.out_stats = sc.take(m);
    q.tag = res_tag;
    RC(conv3x3_c32(net, q, R.res[i], precise, stream));
    st_prev = q.out_stats;
    gn_prev = &R.gn[i];
    ycur = 1 - ycur;
    xres = xnew;
  }
  // conv_final 32 -> 1 on x6 = lrelu(gn(y)) + x5, then relu(prior * fx + delta) / fx
  RC(launch_refine_final(ws.ry[ycur], ws.rx[xres], half_act != 0, st_prev, gn_prev->gamma, gn_prev->beta, inv_count,
                         R.finw, prior, Kl, k_div, 16, m, H, W, out, stream));
  return 0;
}

// Guide part of an IDepthmapRefiner.conv0 (+ bias) for the B reference images of one level: conv0 over
// cat[image, features] with the idepth channel left out (level 0: the image alone), fp32 channels-last.
int run_refiner_guide(b200mvs_net* net, const Refiner& R, int B, int H, int W, const float* guide_feat,
                      const float* image, float* pre, cudaStream_t stream) {
  if (guide_feat == nullptr) return launch_refine_head_image_l0(image, net->head0, B, H, W, pre, stream);
  const size_t P = (size_t)H * W;
  ConvParams p;
  p.n_img = B;
  p.Hi = p.Ho = H;
  p.Wi = p.Wo = W;
  p.feat.ptr = guide_feat;
  p.feat.mode = FEAT_RAW;
  p.feat.img_div = 1;
  p.extra.n = 3;   // the fourth planar slot (idepth * fx) stays zero
  for (int e = 0; e < 3; ++e) {
    p.extra.ptr[e] = image + e * P;
    p.extra.img_stride[e] = 3 * (long long)P;
    p.extra.img_div[e] = 1;
  }
  p.w = R.conv0.w;
  p.bias = R.conv0.bias;
  p.dil = 1;
  p.out = pre;
  return launch_conv3x3_tc(p, R.conv0.w16s, true, stream);
}

// FeatureNetwork.forward (multi_view_stereonet.py:109-129) on images [img0, img0 + cnt) of the (1+V)*B image
// arrays: conv0 reads `cnt` planar 3-channel images at `planar`; `tail_stats[i]` is layer i's statistics block
// for all images.
int run_featnet(b200mvs_net* net, const Levels& L, int img0, int cnt, const float* planar, double* const* tail_stats,
                float* final_out, long long final_stride, cudaStream_t stream) {
  Workspace& ws = net->cur->ws;
  const int h4 = L.h[4], w4 = L.w[4];
  const size_t P4 = L.px[4];
  if (net->use_tensor_cores) {
    // conv0..conv3: 5x5 stride 2 on the tensor cores (split fp16), multi_view_stereonet.py:113-116
    RC(launch_conv5x5s2_c3_tc(planar, net->feat_conv[0].w16s, cnt, L.h[0], L.w[0], ws.f1 + (size_t)img0 * L.px[1] * kC,
                              stream));
    const float* src[3] = {ws.f1, ws.f2, ws.f3};
    float* dst[3] = {ws.f2, ws.f3, ws.l4x[0]};
    for (int i = 0; i < 3; ++i)
      RC(launch_conv5x5s2_c32_tc(src[i] + (size_t)img0 * L.px[i + 1] * kC, net->feat_conv[i + 1].w16s, cnt, L.h[i + 1],
                                 L.w[i + 1], dst[i] + (size_t)img0 * L.px[i + 2] * kC, stream));
  } else {
    {
      ConvParams p;
      p.Hi = L.h[0];
      p.Wi = L.w[0];
      p.Ho = L.h[1];
      p.Wo = L.w[1];
      p.extra.n = 3;
      p.w = net->feat_conv[0].w;
      for (int e = 0; e < 3; ++e) {
        p.extra.ptr[e] = planar + e * L.px[0];
        p.extra.img_stride[e] = 3 * (long long)L.px[0];
      }
      p.n_img = cnt;
      p.out = ws.f1 + (size_t)img0 * L.px[1] * kC;
      RC(launch_conv(CONV_5x5_S2, 32, p, stream));
    }
    const float* src[3] = {ws.f1, ws.f2, ws.f3};
    float* dst[3] = {ws.f2, ws.f3, ws.l4x[0]};
    for (int i = 0; i < 3; ++i) {
      ConvParams p;
      p.n_img = cnt;
      p.Hi = L.h[i + 1];
      p.Wi = L.w[i + 1];
      p.Ho = L.h[i + 2];
      p.Wo = L.w[i + 2];
      p.feat.ptr = src[i] + (size_t)img0 * L.px[i + 1] * kC;
      p.feat.mode = FEAT_RAW;
      p.w = net->feat_conv[i + 1].w;
      p.out = dst[i] + (size_t)img0 * L.px[i + 2] * kC;
      RC(launch_conv(CONV_5x5_S2, 32, p, stream));
    }
  }
  if (net->use_tensor_cores && net->l4_chain && !net->keep_stages && l4_tail_supported(h4, w4) &&
      net->feat_res[0].w16s != nullptr && net->feat_final.w16s != nullptr) {
    // six residual blocks + conv_final at level 4 (multi_view_stereonet.py:119-127) in one cluster kernel per image
    L4TailArgs ta;
    ta.x0 = ws.l4x[0] + (size_t)img0 * P4 * kC;
    ta.out = final_out;
    ta.out_stride = final_stride != 0 ? final_stride : (long long)P4 * kC;
    for (int i = 0; i < 6; ++i) {
      ta.w[i] = net->feat_res[i].w16s;
      ta.bias[i] = net->feat_res[i].bias;
      ta.gamma[i] = net->feat_gn[i].gamma;
      ta.beta[i] = net->feat_gn[i].beta;
    }
    ta.w[6] = net->feat_final.w16s;
    ta.bias[6] = net->feat_final.bias;
    ta.n = cnt;
    ta.rows = h4;
    ta.cols = w4;
    RC(launch_l4_tail(ta, stream));
  } else {
    // six residual blocks + conv_final at level 4 (multi_view_stereonet.py:119-127)
    const double inv_count = 1.0 / (8.0 * (double)P4);
    const size_t ioff = (size_t)img0 * P4 * kC;
    double* st_prev = nullptr;
    int xcur = 0, ycur = 0;
    for (int i = 0; i <= 6; ++i) {
      ConvParams p;
      p.n_img = cnt;
      p.Hi = p.Ho = h4;
      p.Wi = p.Wo = w4;
      if (i == 0) {
        p.feat.ptr = ws.l4x[0] + ioff;
        p.feat.mode = FEAT_RAW;
      } else {
        p.feat.ptr = ws.l4y[ycur] + ioff;
        p.feat.mode = FEAT_GN_RES;
        p.feat.stats = st_prev;
        p.feat.gamma = net->feat_gn[i - 1].gamma;
        p.feat.beta = net->feat_gn[i - 1].beta;
        p.feat.inv_count = inv_count;
        p.feat.resid = ws.l4x[xcur] + ioff;
        if (i < 6) p.feat.x_out = ws.l4x[1 - xcur] + ioff;
      }
      if (i < 6) {
        p.w = net->feat_res[i].w;
        p.out = ws.l4y[i == 0 ? 0 : 1 - ycur] + ioff;
        p.out_stats = tail_stats[i] + (size_t)img0 * 2 * kGroups;
      } else {
        p.w = net->feat_final.w;
        p.bias = net->feat_final.bias;
        p.out = final_out;
        p.out_img_stride = final_stride;
      }
      RC(conv3x3_c32(net, p, i < 6 ? net->feat_res[i] : net->feat_final, true, stream));
      if (i > 0) {
        ycur = 1 - ycur;
        if (i < 6) xcur = 1 - xcur;
      }
      st_prev = p.out_stats;
    }
  }
  return 0;
}

// Shape and pointer checks shared by both forward entries (before anything is staged or allocated).
int validate_call(const b200mvs_shape& s, const float* const* left_pyr, const float* const* K_pyr,
                  const float* const* Ts, const float* const* right_l0, const float* const* right_l4) {
  if (s.batch < 1 || s.views < 1 || s.views > kMaxViews || s.rows < 16 || s.cols < 16 ||
      s.num_idepth_samples < 2 || s.num_idepth_samples > 4096) {
    set_error("b200mvs_forward: bad shape (need batch>=1, 1<=views<=16, rows,cols>=16, 2<=D<=4096)");
    return B200MVS_EINVAL;
  }
  if (left_pyr == nullptr || K_pyr == nullptr || Ts == nullptr || right_l0 == nullptr || right_l4 == nullptr) {
    set_error("b200mvs_forward: null argument");
    return B200MVS_EINVAL;
  }
  for (int l = 0; l < 5; ++l)
    if (left_pyr[l] == nullptr || K_pyr[l] == nullptr) {
      // the reference asserts len(K_pyr) == len(left_image_pyr) == 5 (multi_view_stereonet.py:548-549)
      set_error("b200mvs_forward: left_image_pyr and K_pyr need 5 levels");
      return B200MVS_EINVAL;
    }
  for (int v = 0; v < s.views; ++v)
    if (Ts[v] == nullptr || right_l0[v] == nullptr || right_l4[v] == nullptr) {
      set_error("b200mvs_forward: missing per-view input");
      return B200MVS_EINVAL;
    }
  return 0;
}

int forward_impl(b200mvs_net* net, Lane& lane, bool sweep_on_own_stream, const b200mvs_shape& s,
                 const float* const* left_pyr, const float* const* K_pyr, const float* const* Ts,
                 const float* const* right_l0, const float* const* right_l4, float* const* out_idepth,
                 float* const* out_raw, uint8_t* const* out_mask, cudaStream_t stream,
                 const cudaEvent_t* uploaded = nullptr, cudaEvent_t coarse_done = nullptr) {
  net->cur = &lane;
  net->precise_now = net->precise_refiners == 1 || (net->precise_refiners == 2 && s.num_idepth_samples < 48);
  // `coarse_done` (b200mvs_forward_host): recorded once the idepth maps of levels 1-4 are final, so that their
  // download can run next to the level-0 refiner.
  // `uploaded` (b200mvs_forward_host): events after which [0] K/T, [1] right level 0, [2] right level 4, [3] left
  // level 0, [4] left levels 1-4 are resident; null = all inputs already resident.
  auto wait_upload = [&](int which, cudaStream_t on) -> int {
    if (uploaded != nullptr) B200MVS_CUDA_OK(cudaStreamWaitEvent(on, uploaded[which], 0));
    return 0;
  };
  RC(validate_call(s, left_pyr, K_pyr, Ts, right_l0, right_l4));
  B200MVS_CUDA_OK(cudaSetDevice(net->device));
  RC(check_sweep_watchdog(net));   // (of an earlier asynchronous forward)
  RC(ensure_workspace(net, s));
  Workspace& ws = net->cur->ws;
  const Levels L = levels_of(s);
  const int B = s.batch, V = s.views, D = s.num_idepth_samples;
  const int n = B * V, NI = B + n;
  const int h4 = L.h[4], w4 = L.w[4];
  const size_t P4 = L.px[4];
  g_probe = net->probe.tag != 0 ? &net->probe : nullptr;
  // Debug hook (B200MVS_STAGE_PROFILE=1): events at the stage boundaries of the main stream, printed to stderr
  // after a synchronise.  Perturbs the pipeline slightly (an event between two kernels ends their PDL overlap).
  static const bool sprof_env = getenv("B200MVS_STAGE_PROFILE") != nullptr;
  const bool sprof = sprof_env || net->stage_profile;
  std::vector<std::pair<const char*, cudaEvent_t>> marks;
  auto mark = [&](const char* what) {
    if (lane_trace_enabled() && sweep_on_own_stream) {
      cudaEvent_t e;
      cudaEventCreate(&e);
      cudaEventRecord(e, stream);
      g_trace.push_back({(int)(&lane - net->lanes), what, e});
    }
    if (!sprof) return;
    cudaEvent_t e;
    cudaEventCreate(&e);
    cudaEventRecord(e, stream);
    marks.emplace_back(what, e);
  };
  mark("begin");

  B200MVS_CUDA_OK(cudaMemsetAsync(ws.stats, 0, ws.stats_count * sizeof(double), stream));
  StatsCursor sc{ws.stats, 0, ws.stats_count};

  ViewPtrs Tv{}, R0{}, R4{};
  Tv.views = R0.views = R4.views = V;
  for (int v = 0; v < V; ++v) {
    Tv.p[v] = Ts[v];
    R0.p[v] = right_l0[v];
    R4.p[v] = right_l4[v];
  }

  // The left feature network needs nothing from the comparison views: it runs on the side stream while the main
  // stream does geometry -> warp -> right feature network -> depth-sweep recurrence (which occupies one
  // cluster's worth of SMs per view and leaves the rest of the chip idle at small batch).
  double* tail_stats[6];
  for (int i = 0; i < 6; ++i) tail_stats[i] = sc.take(NI);
  const bool overlap = net->overlap && net->cur->side != nullptr;
  cudaStream_t left_stream = overlap ? net->cur->side : stream;
  if (overlap) {
    B200MVS_CUDA_OK(cudaEventRecord(net->cur->ev_fork, stream));
    B200MVS_CUDA_OK(cudaStreamWaitEvent(net->cur->side, net->cur->ev_fork, 0));
  }
  // 1. geometry
  RC(wait_upload(0, stream));
  RC(launch_geometry(Tv, K_pyr[0], K_pyr[4], B, D, h4, w4, ws.geo, stream));

  // 1b. image half of FeatureRefiner.conv0 for every hypothesis (needs only the homographies and the 1/16-scale
  //     comparison images): first thing on the side stream, the recurrence waits for it
  // Which sweep kernel: the cluster kernel (recurrence.cu) runs one chain per cluster, recurrence_max_clusters (7 on
  // B200) of them at a time, at ~9 us per dependent step; the wide kernel (sweep_wide.cu) takes every chain of the call
  // in one co-resident grid at ~25-40 us per step.  Measured at 512x640 / 64 hypotheses (cluster -> wide): 8 chains
  // 5.03 -> 5.25 ms, 12 chains 4.31 -> 4.39, 16 chains 6.75 -> 6.50 and 9.64 -> 9.31, 32 chains 9.95 -> 8.97, 64
  // chains 19.4 -> 17.4: the wide kernel from the third round of clusters on.  Never next to another lane's sweep
  // (two cooperative grids in flight could each hold SMs the other waits for).
  const bool cluster_ok = recurrence_supported(h4, w4, nullptr, nullptr);
  const bool wide_ok = sweep_wide_supported(h4, w4) && !sweep_on_own_stream;
  bool cluster_sweep = false, wide_sweep = false;
  if (net->use_tensor_cores && net->sweep_mode != 2) {
    if (net->sweep_mode == 1 && wide_ok) {
      wide_sweep = true;
    } else if (cluster_ok) {
      const int maxc = recurrence_max_clusters(h4, w4);
      wide_sweep = wide_ok && net->sweep_mode == 0 && maxc > 0 && n > 2 * maxc && net->rec_prof == nullptr;
      cluster_sweep = !wide_sweep;
    } else {
      wide_sweep = wide_ok;
    }
  }
  const bool persistent = cluster_sweep || wide_sweep;
  if (persistent) {
    if (overlap) {
      B200MVS_CUDA_OK(cudaEventRecord(net->cur->ev_geo, stream));
      B200MVS_CUDA_OK(cudaStreamWaitEvent(net->cur->side, net->cur->ev_geo, 0));
    }
    RC(wait_upload(2, left_stream));
    RC(launch_image_conv(ws.geo.H, R4, net->fr_conv0.w + (size_t)4 * 9 * 8 * 32, net->fr_conv0.bias, n, D, h4, w4,
                         ws.imgconv, left_stream, /*oct_major=*/wide_sweep));
    RC(launch_gather_plan(ws.geo.Hinc, n, D, h4, w4, ws.rec_plan, ws.rec_flags, left_stream));
    if (overlap) B200MVS_CUDA_OK(cudaEventRecord(net->cur->ev_imgconv, net->cur->side));
  }

  // 2. full-resolution warp of every comparison image by the idepth-0 homography
  //    (multi_view_stereonet.py:254-258)
  RC(wait_upload(1, stream));
  RC(launch_warp_planar(ws.geo.H0, 9, R0, n, 3, L.h[0], L.w[0], true, ws.warped0, ws.l0mask, stream));

  // 3b. FeatureNetwork on the B*V warped right images (shared weights, :507)
  //     conv_final writes hypothesis 0 of every view's feature volume directly (:261, 278)
  RC(run_featnet(net, L, B, n, ws.warped0, tail_stats, ws.vol, (long long)D * (long long)P4 * kC, stream));
  if (overlap) B200MVS_CUDA_OK(cudaEventRecord(net->cur->ev_right, stream));

  mark("geometry+warp+right featnet");
  // 3a. FeatureNetwork on the B left images (multi_view_stereonet.py:552): side stream, needed by the cost volume
  //     (enqueued after the critical path's launches)
  //     and not started before the right feature network is through: its 1-CTA-per-SM kernels would take SMs
  //     from the critical path, while nothing needs the left features before the sweep (~0.86 ms) has finished
  if (overlap && net->left_late) B200MVS_CUDA_OK(cudaStreamWaitEvent(net->cur->side, net->cur->ev_right, 0));
  RC(wait_upload(3, left_stream));
  RC(run_featnet(net, L, 0, B, left_pyr[0], tail_stats, ws.feat4, 0, left_stream));
  if (overlap) B200MVS_CUDA_OK(cudaEventRecord(net->cur->ev_left, net->cur->side));

  // 3c. every refiner's conv0 is linear in its inputs and all but one of them (image, left features) are known
  //     now: their part is computed here, next to the depth sweep; the idepth channel is added when the coarser
  //     level has produced it (run_refiner)
  bool use_pre[5] = {false, false, false, false, false};
  if (net->conv0_precompute && net->use_tensor_cores) {
    // not before the depth sweep is ready to start: its cluster needs 11 free SMs in one GPC, and a large grid
    // already resident there would hold it back (measured: +29 us on the sweep); smallest level first
    if (overlap) B200MVS_CUDA_OK(cudaStreamWaitEvent(net->cur->side, net->cur->ev_right, 0));
    RC(wait_upload(4, left_stream));
    for (int l = 4; l >= 0; --l) {
      if (!s.do_refiners[l]) continue;
      const float* guide = (l == 0) ? nullptr : (l == 1 ? ws.f1 : (l == 2 ? ws.f2 : (l == 3 ? ws.f3 : ws.feat4)));
      RC(run_refiner_guide(net, net->refiner[l], B, L.h[l], L.w[l], guide, left_pyr[l], ws.pre[l], left_stream));
      use_pre[l] = true;
    }
    if (overlap) B200MVS_CUDA_OK(cudaEventRecord(net->cur->ev_pre, net->cur->side));
  }

  // Output buffers, and the mask volumes: they depend on the cameras only (a voxel is masked where the views'
  // homographies leave the image, :293-298, voted over the views :623-627, then upsampled level by level :389-396), so
  // their whole chain runs on the side stream next to the depth sweep.  (Run next to the refiners, its large grids
  // held back the small kernels of refiners 3 and 2: 32 us per forward at batch 1, 330 us at batch 8.)
  float* idepth_l[5];
  float* prior_l[5];
  uint8_t* mask_l[5];
  int lowest_mask = 5;
  for (int l = 0; l < 5; ++l) {
    idepth_l[l] = out_idepth != nullptr && out_idepth[l] != nullptr ? out_idepth[l] : ws.idepth[l];
    prior_l[l] = out_raw != nullptr && out_raw[l] != nullptr ? out_raw[l] : ws.prior[l];
    mask_l[l] = out_mask != nullptr && out_mask[l] != nullptr ? out_mask[l] : ws.mask[l];
    if (out_mask != nullptr && out_mask[l] != nullptr && l < lowest_mask) lowest_mask = l;
  }
  const bool early_masks = overlap;
  if (early_masks) {
    RC(launch_mask_vote(ws.geo.H, B, V, D, h4, w4, mask_l[4], net->cur->side));
    for (int l = 3; l >= lowest_mask; --l)
      RC(launch_upsample_mask(mask_l[l + 1], (long long)B * D, L.h[l + 1], L.w[l + 1], L.h[l], L.w[l], mask_l[l],
                              net->cur->side));
    B200MVS_CUDA_OK(cudaEventRecord(net->cur->ev_mask_out, net->cur->side));
  }

  // 5. the depth-sweep recurrence (multi_view_stereonet.py:279-290)
  RC(wait_upload(2, stream));
  if (overlap && persistent) B200MVS_CUDA_OK(cudaStreamWaitEvent(stream, net->cur->ev_imgconv, 0));
  if (persistent) {
    RecurrenceArgs ra;
    ra.imgconv = ws.imgconv;
    ra.plan = ws.rec_plan;
    ra.flags = ws.rec_flags;
    ra.vol = ws.vol;
    ra.geo = ws.geo;
    ra.right_l4 = R4;
    ra.w16 = net->fr_w16;
    ra.bias0 = net->fr_conv0.bias;
    ra.bias1 = net->fr_res0.bias;
    ra.bias2 = net->fr_final.bias;
    ra.gamma0 = net->fr_gn0.gamma;
    ra.beta0 = net->fr_gn0.beta;
    ra.gamma1 = net->fr_gn1.gamma;
    ra.beta1 = net->fr_gn1.beta;
    ra.n = n;
    ra.D = D;
    ra.rows = h4;
    ra.cols = w4;
    ra.prof = net->rec_prof;
    ra.debug = net->rec_debug;
    if (wide_sweep) {
      probe_before(TAG_RECURRENCE, stream);
      if (net->wide_abort_host == nullptr) {
        B200MVS_CUDA_OK(cudaHostAlloc(reinterpret_cast<void**>(&net->wide_abort_host), sizeof(unsigned), cudaHostAllocDefault));
        *net->wide_abort_host = 0;
      }
      ws.wide.abort_host = net->wide_abort_host;
      RC(launch_sweep_wide(ra, ws.wide, stream));
      probe_after(TAG_RECURRENCE, stream);
    } else if (sweep_on_own_stream) {
      // several lanes in flight: the sweep's cluster launch goes through the lane's high-priority stream, so that
      // its clusters are placed as soon as a GPC can take them, ahead of the other lanes' pending thread blocks
      B200MVS_CUDA_OK(cudaEventRecord(lane.ev_rec_in, stream));
      B200MVS_CUDA_OK(cudaStreamWaitEvent(lane.rec, lane.ev_rec_in, 0));
      RC(launch_recurrence(ra, lane.rec));
      B200MVS_CUDA_OK(cudaEventRecord(lane.ev_rec_out, lane.rec));
      B200MVS_CUDA_OK(cudaStreamWaitEvent(stream, lane.ev_rec_out, 0));
    } else {
      probe_before(TAG_RECURRENCE, stream);
      RC(launch_recurrence(ra, stream));
      probe_after(TAG_RECURRENCE, stream);
    }
  } else {
    const double inv_count = 1.0 / (8.0 * (double)P4);
    for (int step = 1; step < D; ++step) {
      RC(launch_step_warp(ws.vol, ws.geo, R4, n, D, step, h4, w4, ws.wf, ws.wimg, stream));
      ConvParams a;  // conv0 on cat([warped image, warped features])
      a.n_img = n;
      a.Hi = a.Ho = h4;
      a.Wi = a.Wo = w4;
      a.feat.ptr = ws.wf;
      a.feat.mode = FEAT_RAW;
      a.extra.n = 3;
      for (int e = 0; e < 3; ++e) {
        a.extra.ptr[e] = ws.wimg + e * P4;
        a.extra.img_stride[e] = 3 * (long long)P4;
      }
      a.w = net->fr_conv0.w;
      a.bias = net->fr_conv0.bias;
      a.out = ws.sy0;
      a.out_stats = sc.take(n);
      RC(conv3x3_c32(net, a, net->fr_conv0, true, stream));

      ConvParams b;  // res0.conv1 on x0 = lrelu(gn0(y0))
      b.n_img = n;
      b.Hi = b.Ho = h4;
      b.Wi = b.Wo = w4;
      b.feat.ptr = ws.sy0;
      b.feat.mode = FEAT_GN;
      b.feat.stats = a.out_stats;
      b.feat.gamma = net->fr_gn0.gamma;
      b.feat.beta = net->fr_gn0.beta;
      b.feat.inv_count = inv_count;
      b.feat.x_out = ws.sx0;
      b.w = net->fr_res0.w;
      b.bias = net->fr_res0.bias;
      b.out = ws.sy1;
      b.out_stats = sc.take(n);
      RC(conv3x3_c32(net, b, net->fr_res0, true, stream));

      ConvParams c;  // conv_final on x1 = lrelu(gn1(y1)) + x0 ; features_d = warped + delta
      c.n_img = n;
      c.Hi = c.Ho = h4;
      c.Wi = c.Wo = w4;
      c.feat.ptr = ws.sy1;
      c.feat.mode = FEAT_GN_RES;
      c.feat.stats = b.out_stats;
      c.feat.gamma = net->fr_gn1.gamma;
      c.feat.beta = net->fr_gn1.beta;
      c.feat.inv_count = inv_count;
      c.feat.resid = ws.sx0;
      c.w = net->fr_final.w;
      c.bias = net->fr_final.bias;
      c.add_src = ws.wf;
      // hypothesis `step` of image i lives at vol + (i * D + step) * P4 * 32
      c.out = ws.vol + (size_t)step * P4 * kC;
      c.out_img_stride = (long long)D * P4 * kC;
      RC(conv3x3_c32(net, c, net->fr_final, true, stream));
    }
  }

  mark("recurrence");
  // 6. cost volume |L - R| with invalid voxels zeroed (multi_view_stereonet.py:586-592)
  if (overlap) B200MVS_CUDA_OK(cudaStreamWaitEvent(stream, net->cur->ev_left, 0));
  float* cost = net->keep_stages ? ws.cost : ws.vol;
  RC(launch_cost(ws.feat4, ws.vol, ws.geo.H, n, V, D, h4, w4, cost, ws.mask_views, stream));

  // 7. CostVolumeFilter (five Conv3d, :341-353) or the channel norm (:598)
  bool softargmin_done = false;
  if (s.do_cost_volume_filter) {
    const double inv_count = 1.0 / (8.0 * (double)D * (double)P4);
    float* bufs[2] = {ws.cvfA, cost};
    const float* src = cost;
    double* st_prev = nullptr;
    for (int i = 0; i < 5; ++i) {
      ConvParams p;
      p.n_img = n;
      p.Di = p.Do = D;
      p.Hi = p.Ho = h4;
      p.Wi = p.Wo = w4;
      p.feat.ptr = src;
      if (i == 0) {
        p.feat.mode = FEAT_RAW;
      } else {
        p.feat.mode = FEAT_GN;
        p.feat.stats = st_prev;
        p.feat.gamma = net->cvf_gn[i - 1].gamma;
        p.feat.beta = net->cvf_gn[i - 1].beta;
        p.feat.inv_count = inv_count;
      }
      p.w = net->cvf[i].w;
      p.bias = net->cvf[i].bias;
      if (i < 4) {
        p.out = bufs[i & 1];
        p.out_stats = sc.take(n);
        p.tag = TAG_CVF_CONV32;
        if (net->use_tensor_cores && cvf_tc_supported(h4, w4)) {
          CvfArgs ca;
          ca.in = p.feat.ptr;
          ca.mode = p.feat.mode;
          ca.stats = p.feat.stats;
          ca.gamma = p.feat.gamma;
          ca.beta = p.feat.beta;
          ca.inv_count = p.feat.inv_count;
          ca.w16 = net->cvf[i].w16;
          ca.bias = p.bias;
          ca.out = p.out;
          ca.out_stats = p.out_stats;
          ca.n = n;
          ca.D = D;
          ca.h = h4;
          ca.w = w4;
          ca.tag = p.tag;
          RC(launch_cvf_tc(ca, stream));
        } else {
          RC(launch_conv(CONV_3x3x3, 32, p, stream));
        }
        src = p.out;
        st_prev = p.out_stats;
      } else if (cvf_final_supported(D)) {
        // conv4 (32 -> 1) fused with the soft-argmin (:350-352, 602); the idle ping-pong buffer holds the partials
        RC(launch_cvf_final(src, st_prev, net->cvf_gn[3].gamma, net->cvf_gn[3].beta, inv_count, net->cvf_finw,
                            ws.geo.samples, n, D, h4, w4, bufs[0], ws.cost1, ws.raw_views, stream));
        softargmin_done = true;
      } else {
        p.out = ws.cost1;
        RC(launch_conv(CONV_3x3x3, 1, p, stream));
      }
    }
  } else {
    RC(launch_cost_norm(cost, (long long)n * D * P4, ws.cost1, stream));
  }

  // 8. soft-argmin (:602)
  if (!softargmin_done) RC(launch_softargmin(ws.cost1, ws.geo.samples, n, D, (int)P4, ws.raw_views, stream));

  mark("cost+cvf+softargmin");
  // 9. level-4 refiner per view (:605-613)
  RC(wait_upload(4, stream));
  if (overlap && net->conv0_precompute && net->use_tensor_cores) B200MVS_CUDA_OK(cudaStreamWaitEvent(stream, net->cur->ev_pre, 0));
  if (s.do_refiners[4]) {
    RC(run_refiner(net, net->refiner[4], sc, n, h4, w4, ws.feat4, V, left_pyr[4], V, ws.raw_views, K_pyr[4], V,
                   ws.refined_views, TAG_NONE, stream, use_pre[4] ? ws.pre[4] : nullptr));
  }

  // 10. baseline un-normalisation, mean over views, mask vote (:616-627)
  RC(launch_view_reduce(ws.raw_views, ws.refined_views, ws.mask_views, ws.geo.baseline, B, V, D, (int)P4,
                        !s.do_refiners[4], prior_l[4], idepth_l[4], early_masks ? nullptr : mask_l[4], stream));

  mark("refiner4+view_reduce");
  // 11. coarse-to-fine: bilinear prior, mask upsample, guided refinement (:629-682).  (The mask volumes were
  //     produced next to the depth sweep unless the side stream is off.)
  if (!early_masks)
    for (int l = 3; l >= lowest_mask; --l)
      RC(launch_upsample_mask(mask_l[l + 1], (long long)B * D, L.h[l + 1], L.w[l + 1], L.h[l], L.w[l], mask_l[l], stream));
  for (int l = 3; l >= 0; --l) {
    const bool fused_up = s.do_refiners[l] && use_pre[l];   // the refiner's head upsamples its own prior
    if (!fused_up) RC(launch_upsample_f32(idepth_l[l + 1], B, L.h[l + 1], L.w[l + 1], L.h[l], L.w[l], prior_l[l], stream));
    if (s.do_refiners[l]) {
      const float* guide = (l == 0) ? nullptr : (l == 1 ? ws.f1 : (l == 2 ? ws.f2 : ws.f3));
      RC(run_refiner(net, net->refiner[l], sc, B, L.h[l], L.w[l], guide, 1, left_pyr[l], 1, prior_l[l], K_pyr[l], 1,
                     idepth_l[l], l == 0 ? TAG_REFINE_CONV32_L0 : TAG_NONE, stream, use_pre[l] ? ws.pre[l] : nullptr,
                     fused_up ? idepth_l[l + 1] : nullptr, L.h[l + 1], L.w[l + 1]));
    } else {
      B200MVS_CUDA_OK(cudaMemcpyAsync(idepth_l[l], prior_l[l], (size_t)B * L.px[l] * sizeof(float),
                                      cudaMemcpyDeviceToDevice, stream));
    }
    static const char* names[4] = {"upsample+refiner0", "upsample+refiner1", "upsample+refiner2", "upsample+refiner3"};
    mark(names[l]);
    if (l == 1 && coarse_done != nullptr) B200MVS_CUDA_OK(cudaEventRecord(coarse_done, stream));
  }
  if (early_masks) B200MVS_CUDA_OK(cudaStreamWaitEvent(stream, net->cur->ev_mask_out, 0));
  if (sprof) {
    mark("join mask chain");
    cudaEventSynchronize(marks.back().second);
    float total = 0.f;
    cudaEventElapsedTime(&total, marks.front().second, marks.back().second);
    if (sprof_env) fprintf(stderr, "stage profile (us):");
    net->last_stage_profile.clear();
    for (size_t i = 1; i < marks.size(); ++i) {
      float ms = 0.f;
      cudaEventElapsedTime(&ms, marks[i - 1].second, marks[i].second);
      if (sprof_env) fprintf(stderr, " [%s] %.0f", marks[i].first, ms * 1e3f);
      net->last_stage_profile += std::string(marks[i].first) + "=" + std::to_string(ms * 1e3f) + ";";
    }
    net->last_stage_profile += "total=" + std::to_string(total * 1e3f);
    if (sprof_env) fprintf(stderr, " | total %.0f\n", total * 1e3f);
    for (auto& m : marks) cudaEventDestroy(m.second);
  }
  if (sc.used > sc.cap) {
    set_error("internal: GroupNorm statistics arena overflow");
    return B200MVS_EINVAL;
  }
  return 0;
}

// How many lanes a call runs as: 1 unless the depth sweep of its B * V (image group, view) pairs needs more than one
// round of clusters -- then enough lanes that the first ones are through their sweep while the others still hold
// clusters (whole image groups per lane, at most lanes_max).
// lanes_max = 0 (default) decides by measurement: two lanes exactly when the sweep's last round would hold one or two
// clusters (batch 8 with one view on B200: 7 clusters are co-resident, the 8th ran alone for 0.6 ms; as two lanes of
// four groups the lonely cluster overlaps the other lane's cost filter and refiners: 5.15 -> 5.01 ms).  With fuller
// rounds two lanes lose (batch 8 with 2 / 4 views: 6.76 -> 7.42 ms, 9.94 -> 11.2 ms; batch 16: 9.64 -> 10.2 ms).
int plan_lanes(const b200mvs_net* net, const b200mvs_shape& s) {
  const bool automatic = net->lanes_max == 0;
  if ((!automatic && net->lanes_max <= 1) || !net->use_tensor_cores || !net->overlap || net->keep_stages ||
      net->stage_profile || net->probe.tag != 0 || net->rec_prof != nullptr || s.batch < 2 || net->sweep_mode != 0)
    return 1;   // (the wide sweep is a cooperative grid: never two of them in flight)
  static const bool sprof_env = getenv("B200MVS_STAGE_PROFILE") != nullptr;
  if (sprof_env) return 1;
  const Levels L = levels_of(s);
  if (!recurrence_supported(L.h[4], L.w[4], nullptr, nullptr)) return 1;
  const int maxc = recurrence_max_clusters(L.h[4], L.w[4]);
  const int pairs = s.batch * s.views;
  if (maxc < 1 || pairs <= maxc) return 1;
  if (automatic) return (pairs - maxc <= 2 && pairs <= 2 * maxc) ? 2 : 1;
  int want = (pairs + maxc - 1) / maxc;
  if (want > net->lanes_max) want = net->lanes_max;
  if (want > s.batch) want = s.batch;
  const int per = (s.batch + want - 1) / want;   // image groups per lane
  return (s.batch + per - 1) / per;
}

}  // namespace

// ------------------------------------------------------------------------------------------------
// C ABI
// ------------------------------------------------------------------------------------------------
extern "C" {

B200MVS_API const char* b200mvs_last_error(void) { return g_error.c_str(); }
B200MVS_API const char* b200mvs_version(void) { return "b200mvs 0.1 (sm_100a)"; }

B200MVS_API int b200mvs_create(int device, int num_tensors, const char* const* names, const float* const* data,
                   const int64_t* numels, b200mvs_net** out) {
  if (out == nullptr || names == nullptr || data == nullptr || numels == nullptr) {
    set_error("b200mvs_create: null argument");
    return B200MVS_EINVAL;
  }
  *out = nullptr;
  int count = 0;
  if (cudaGetDeviceCount(&count) != cudaSuccess || count == 0) {
    set_error("b200mvs_create: no CUDA device visible; this library has no CPU path");
    return B200MVS_ECUDA;
  }
  B200MVS_CUDA_OK(cudaSetDevice(device));
  cudaDeviceProp prop;
  B200MVS_CUDA_OK(cudaGetDeviceProperties(&prop, device));
  if (prop.major != 10) {
    set_error(std::string("b200mvs_create: built for sm_100a, device is sm_") + std::to_string(prop.major) +
              std::to_string(prop.minor));
    return B200MVS_ECUDA;
  }
  StateDict sd;
  for (int i = 0; i < num_tensors; ++i) sd.t[names[i]] = {data[i], numels[i]};
  b200mvs_net* net = new b200mvs_net();
  net->device = device;
  if (const char* env = getenv("B200MVS_LANES")) net->lanes_max = atoi(env) < 0 ? 0 : (atoi(env) > kMaxLanes ? kMaxLanes : atoi(env));
  int rc = build_weights(net, sd);
  if (rc == 0) {
    bool ok = true;
    auto ev = [&](cudaEvent_t* e) { ok = ok && cudaEventCreateWithFlags(e, cudaEventDisableTiming) == cudaSuccess; };
    int prio_low = 0, prio_high = 0;
    cudaDeviceGetStreamPriorityRange(&prio_low, &prio_high);
    for (Lane& ln : net->lanes) {
      ok = ok && cudaStreamCreateWithFlags(&ln.main, cudaStreamNonBlocking) == cudaSuccess;
      ok = ok && cudaStreamCreateWithFlags(&ln.side, cudaStreamNonBlocking) == cudaSuccess;
      ok = ok && cudaStreamCreateWithPriority(&ln.rec, cudaStreamNonBlocking, prio_high) == cudaSuccess;
      for (cudaEvent_t* e : {&ln.ev_fork, &ln.ev_geo, &ln.ev_imgconv, &ln.ev_left, &ln.ev_pre, &ln.ev_right, &ln.ev_mask_in,
                             &ln.ev_mask_out, &ln.ev_rec_in, &ln.ev_rec_out, &ln.ev_end})
        ev(e);
    }
    ev(&net->ev_coarse);
    ev(&net->ev_begin);
    for (cudaEvent_t& e : net->ev_up) ev(&e);
    ok = ok && cudaStreamCreateWithFlags(&net->copy_stream, cudaStreamNonBlocking) == cudaSuccess;
    ok = ok && cudaStreamCreateWithFlags(&net->host_stream, cudaStreamNonBlocking) == cudaSuccess;
    if (!ok) {
      set_error("b200mvs_create: could not create the side stream");
      rc = B200MVS_ECUDA;
    }
  }
  if (rc != 0) {
    b200mvs_destroy(net);
    return rc;
  }
  *out = net;
  return 0;
}

B200MVS_API void b200mvs_destroy(b200mvs_net* net) {
  if (net == nullptr) return;
  cudaSetDevice(net->device);
  for (void* p : net->allocs) cudaFree(p);
  for (Lane& ln : net->lanes) {
    if (ln.arena.base != nullptr) cudaFree(ln.arena.base);
    for (cudaEvent_t e : {ln.ev_fork, ln.ev_geo, ln.ev_imgconv, ln.ev_left, ln.ev_pre, ln.ev_right, ln.ev_mask_in,
                          ln.ev_mask_out, ln.ev_rec_in, ln.ev_rec_out, ln.ev_end})
      if (e != nullptr) cudaEventDestroy(e);
    for (cudaStream_t st : {ln.main, ln.side, ln.rec})
      if (st != nullptr) cudaStreamDestroy(st);
  }
  if (net->host_stage != nullptr) cudaFree(net->host_stage);
  for (cudaEvent_t e : net->probe.ev) cudaEventDestroy(e);
  for (cudaEvent_t e : {net->ev_coarse, net->ev_begin})
    if (e != nullptr) cudaEventDestroy(e);
  for (cudaEvent_t e : net->ev_up)
    if (e != nullptr) cudaEventDestroy(e);
  if (net->copy_stream != nullptr) cudaStreamDestroy(net->copy_stream);
  if (net->host_stream != nullptr) cudaStreamDestroy(net->host_stream);
  if (net->pinned_small != nullptr) cudaFreeHost(net->pinned_small);
  if (net->wide_abort_host != nullptr) cudaFreeHost(net->wide_abort_host);
  delete net;
}

B200MVS_API int b200mvs_set_debug(b200mvs_net* net, int keep_stages) {
  if (net == nullptr) return B200MVS_EINVAL;
  if (net->keep_stages != (keep_stages != 0)) {
    net->keep_stages = keep_stages != 0;
    for (Lane& ln : net->lanes) ln.ws_valid = false;  // layout changes
  }
  return 0;
}

B200MVS_API int b200mvs_set_option(b200mvs_net* net, const char* name, int value) {
  if (net == nullptr || name == nullptr) return B200MVS_EINVAL;
  const std::string k(name);
  if (k == "tensor_cores") {
    net->use_tensor_cores = value != 0;
    return 0;
  }
  if (k == "warp_specialized") {
    net->warp_specialized = value != 0;
    return 0;
  }
  if (k == "early_d2h") {
    net->early_d2h = value != 0;
    return 0;
  }
  if (k == "left_late") {
    net->left_late = value != 0;
    return 0;
  }
  if (k == "conv0_precompute") {
    net->conv0_precompute = value != 0;
    return 0;
  }
  if (k == "half_activations") {
    net->half_activations = value != 0;
    return 0;
  }
  if (k == "pdl") {
    set_pdl_enabled(value != 0);
    return 0;
  }
  if (k == "overlap") {
    net->overlap = value != 0;
    return 0;
  }
  if (k == "precise_refiners") {
    net->precise_refiners = value;
    return 0;
  }
  if (k == "lanes") {
    net->lanes_max = value < 0 ? 0 : (value > kMaxLanes ? kMaxLanes : value);
    return 0;
  }
  if (k == "stage_profile") {
    net->stage_profile = value != 0;
    return 0;
  }
  if (k == "prio_main") {
    net->prio_main = value != 0;
    return 0;
  }
  if (k == "l4_chain") {
    net->l4_chain = value != 0;
    return 0;
  }
  if (k == "sweep") {
    net->sweep_mode = value < 0 || value > 2 ? 0 : value;
    return 0;
  }
  if (k == "recurrence_debug") {
    net->rec_debug = value;
    return 0;
  }
  if (k == "recurrence_profile") {
    if (value != 0 && net->rec_prof == nullptr) {
      B200MVS_CUDA_OK(cudaMalloc(reinterpret_cast<void**>(&net->rec_prof), kRecProfWords * sizeof(long long)));
      B200MVS_CUDA_OK(cudaMemset(net->rec_prof, 0, kRecProfWords * sizeof(long long)));
      net->allocs.push_back(net->rec_prof);
    }
    return 0;
  }
  set_error("b200mvs_set_option: unknown option '" + k + "'");
  return B200MVS_EINVAL;
}

B200MVS_API int b200mvs_conv3x3_c32(const float* x, const float* w_oihw_host, const float* bias_host, int32_t n,
                                    int32_t rows, int32_t cols, int32_t dilation, int32_t use_tensor_cores, float* y,
                                    void* stream_) {
  if (x == nullptr || w_oihw_host == nullptr || y == nullptr || n < 1 || rows < 1 || cols < 1 || dilation < 1 ||
      dilation > 8) {
    set_error("b200mvs_conv3x3_c32: bad argument");
    return B200MVS_EINVAL;
  }
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  float* dbias = nullptr;
  void* dw = nullptr;
  int rc = 0;
  ConvParams p;
  p.n_img = n;
  p.Hi = p.Ho = rows;
  p.Wi = p.Wo = cols;
  p.dil = dilation;
  p.feat.ptr = x;
  p.feat.mode = FEAT_RAW;
  p.out = y;
  if (bias_host != nullptr) {
    B200MVS_CUDA_OK(cudaMalloc(reinterpret_cast<void**>(&dbias), 32 * sizeof(float)));
    B200MVS_CUDA_OK(cudaMemcpyAsync(dbias, bias_host, 32 * sizeof(float), cudaMemcpyHostToDevice, stream));
    p.bias = dbias;
  }
  if (use_tensor_cores) {
    std::vector<uint8_t> packed;
    pack_conv3x3_tc_weights(w_oihw_host, 32, true, 0, {}, use_tensor_cores == 2, &packed);
    B200MVS_CUDA_OK(cudaMalloc(&dw, packed.size()));
    B200MVS_CUDA_OK(cudaMemcpy(dw, packed.data(), packed.size(), cudaMemcpyHostToDevice));
    rc = launch_conv3x3_tc(p, static_cast<const uint8_t*>(dw), use_tensor_cores == 2, stream);
  } else {
    std::vector<float> packed((size_t)4 * 9 * 8 * 32);
    for (int chunk = 0; chunk < 4; ++chunk)
      for (int tap = 0; tap < 9; ++tap)
        for (int k = 0; k < 8; ++k)
          for (int o = 0; o < 32; ++o)
            packed[(((size_t)chunk * 9 + tap) * 8 + k) * 32 + o] = w_oihw_host[((size_t)o * 32 + chunk * 8 + k) * 9 + tap];
    B200MVS_CUDA_OK(cudaMalloc(&dw, packed.size() * sizeof(float)));
    B200MVS_CUDA_OK(cudaMemcpy(dw, packed.data(), packed.size() * sizeof(float), cudaMemcpyHostToDevice));
    p.w = static_cast<const float*>(dw);
    rc = launch_conv(CONV_3x3, 32, p, stream);
  }
  cudaStreamSynchronize(stream);
  cudaFree(dw);
  if (dbias != nullptr) cudaFree(dbias);
  return rc;
}

B200MVS_API int b200mvs_probe_select(b200mvs_net* net, const char* kernel_class) {
  if (net == nullptr || kernel_class == nullptr) return B200MVS_EINVAL;
  const std::string k(kernel_class);
  int tag;
  if (k == "none") tag = TAG_NONE;
  else if (k == "refine_conv32_l0") tag = TAG_REFINE_CONV32_L0;
  else if (k == "cvf_conv32") tag = TAG_CVF_CONV32;
  else if (k == "recurrence") tag = TAG_RECURRENCE;
  else {
    set_error("b200mvs_probe_select: unknown kernel class '" + k + "'");
    return B200MVS_EINVAL;
  }
  net->probe.tag = tag;
  net->probe.used = 0;
  return 0;
}

B200MVS_API int b200mvs_probe_read(b200mvs_net* net, double* total_ms, int64_t* launches) {
  if (net == nullptr) return B200MVS_EINVAL;
  double total = 0.0;
  const size_t pairs = net->probe.used / 2;
  for (size_t i = 0; i < pairs; ++i) {
    B200MVS_CUDA_OK(cudaEventSynchronize(net->probe.ev[2 * i + 1]));
    float ms = 0.f;
    B200MVS_CUDA_OK(cudaEventElapsedTime(&ms, net->probe.ev[2 * i], net->probe.ev[2 * i + 1]));
    total += ms;
  }
  if (total_ms != nullptr) *total_ms = total;
  if (launches != nullptr) *launches = (int64_t)pairs;
  net->probe.used = 0;
  return 0;
}

B200MVS_API int64_t b200mvs_last_launch_count(const b200mvs_net* net) { return net == nullptr ? 0 : net->last_launches; }

B200MVS_API const char* b200mvs_last_stage_profile(const b200mvs_net* net) {
  return net == nullptr ? "" : net->last_stage_profile.c_str();
}

B200MVS_API int b200mvs_forward(b200mvs_net* net, const b200mvs_shape* shape, const float* const* left_image_pyr,
                    const float* const* K_pyr, const float* const* T_right_in_lefts,
                    const float* const* right_image_l0, const float* const* right_image_l4,
                    float* const* out_idepth, float* const* out_idepth_raw, uint8_t* const* out_mask,
                    void* stream) {
  if (net == nullptr || shape == nullptr || left_image_pyr == nullptr || K_pyr == nullptr ||
      T_right_in_lefts == nullptr || right_image_l0 == nullptr || right_image_l4 == nullptr) {
    set_error("b200mvs_forward: null argument");
    return B200MVS_EINVAL;
  }
  cudaStream_t caller = static_cast<cudaStream_t>(stream);
  const b200mvs_shape& s = *shape;
  RC(validate_call(s, left_image_pyr, K_pyr, T_right_in_lefts, right_image_l0, right_image_l4));
  g_launches = 0;
  const int lanes = plan_lanes(net, s);
  int rc = 0;
  if (lanes == 1 && net->prio_main && net->overlap) {
    // experiment: the critical path on a high-priority stream of our own, so that the side stream's grids only get
    // the SMs it leaves idle
    Lane& l0 = net->lanes[0];
    B200MVS_CUDA_OK(cudaSetDevice(net->device));
    B200MVS_CUDA_OK(cudaEventRecord(net->ev_begin, caller));
    B200MVS_CUDA_OK(cudaStreamWaitEvent(l0.rec, net->ev_begin, 0));
    rc = forward_impl(net, l0, false, s, left_image_pyr, K_pyr, T_right_in_lefts, right_image_l0, right_image_l4,
                      out_idepth, out_idepth_raw, out_mask, l0.rec);
    B200MVS_CUDA_OK(cudaEventRecord(l0.ev_end, l0.rec));
    B200MVS_CUDA_OK(cudaStreamWaitEvent(caller, l0.ev_end, 0));
  } else if (lanes == 1) {
    rc = forward_impl(net, net->lanes[0], false, s, left_image_pyr, K_pyr, T_right_in_lefts, right_image_l0,
                      right_image_l4, out_idepth, out_idepth_raw, out_mask, caller);
  } else {
    // Lane g owns image groups [g * per, min(B, (g + 1) * per)): every input and output is batch-major, so a lane's
    // tensors are the caller's pointers advanced by whole image groups.
    B200MVS_CUDA_OK(cudaSetDevice(net->device));
    const Levels L = levels_of(s);
    const int per = (s.batch + lanes - 1) / lanes;
    B200MVS_CUDA_OK(cudaEventRecord(net->ev_begin, caller));
    {
      const int maxc = recurrence_max_clusters(L.h[4], L.w[4]);
      const int other = per * s.views < maxc ? per * s.views : maxc;   // clusters another lane can hold at once
      set_reserved_sms(other * 11);
    }
    if (lane_trace_enabled()) {
      cudaEventCreate(&g_trace_origin);
      cudaEventRecord(g_trace_origin, caller);
    }
    for (int g = 0; g < lanes && rc == 0; ++g) {
      Lane& lane = net->lanes[g];
      const size_t b0 = (size_t)g * per;
      b200mvs_shape sub = s;
      sub.batch = (int)((size_t)s.batch - b0 < (size_t)per ? (size_t)s.batch - b0 : (size_t)per);
      const float *lp[5], *kp[5], *tp[kMaxViews], *r0[kMaxViews], *r4[kMaxViews];
      float *oi[5], *orw[5];
      uint8_t* om[5];
      for (int l = 0; l < 5; ++l) {
        lp[l] = left_image_pyr[l] + b0 * 3 * L.px[l];
        kp[l] = K_pyr[l] + b0 * 16;
        oi[l] = out_idepth != nullptr && out_idepth[l] != nullptr ? out_idepth[l] + b0 * L.px[l] : nullptr;
        orw[l] = out_idepth_raw != nullptr && out_idepth_raw[l] != nullptr ? out_idepth_raw[l] + b0 * L.px[l] : nullptr;
        om[l] = out_mask != nullptr && out_mask[l] != nullptr
                    ? out_mask[l] + b0 * (size_t)s.num_idepth_samples * L.px[l] : nullptr;
      }
      for (int v = 0; v < s.views; ++v) {
        tp[v] = T_right_in_lefts[v] + b0 * 16;
        r0[v] = right_image_l0[v] + b0 * 3 * L.px[0];
        r4[v] = right_image_l4[v] + b0 * 3 * L.px[4];
      }
      B200MVS_CUDA_OK(cudaStreamWaitEvent(lane.main, net->ev_begin, 0));
      rc = forward_impl(net, lane, true, sub, lp, kp, tp, r0, r4, out_idepth != nullptr ? oi : nullptr,
                        out_idepth_raw != nullptr ? orw : nullptr, out_mask != nullptr ? om : nullptr, lane.main);
      B200MVS_CUDA_OK(cudaEventRecord(lane.ev_end, lane.main));
      B200MVS_CUDA_OK(cudaStreamWaitEvent(caller, lane.ev_end, 0));
    }
    net->cur = &net->lanes[0];
    set_reserved_sms(0);
    if (lane_trace_enabled()) {
      cudaStreamSynchronize(caller);
      for (int g = 0; g < lanes; ++g) {
        fprintf(stderr, "lane %d (us):", g);
        for (const TraceMark& m : g_trace)
          if (m.lane == g) {
            float ms = 0.f;
            cudaEventElapsedTime(&ms, g_trace_origin, m.ev);
            fprintf(stderr, " %s@%.0f", m.what, ms * 1e3f);
          }
        fprintf(stderr, "\n");
      }
      for (const TraceMark& m : g_trace) cudaEventDestroy(m.ev);
      g_trace.clear();
      cudaEventDestroy(g_trace_origin);
    }
  }
  if (rc == 0) {
    net->last_shape = s;
    net->have_last = true;
    net->last_launches = g_launches;
    net->last_lanes = lanes;
  }
  return rc;
}

B200MVS_API int b200mvs_forward_host(b200mvs_net* net, const b200mvs_shape* shape, const float* const* left_image_pyr,
                         const float* const* K_pyr, const float* const* T_right_in_lefts,
                         const float* const* right_image_l0, const float* const* right_image_l4,
                         float* const* out_idepth, float* const* out_idepth_raw, uint8_t* const* out_mask,
                         int64_t* h2d_bytes, int64_t* d2h_bytes) {
  if (net == nullptr || shape == nullptr) {
    set_error("b200mvs_forward_host: null argument");
    return B200MVS_EINVAL;
  }
  const b200mvs_shape& s = *shape;
  // full validation BEFORE any staging: the pointer arrays are dereferenced and the staging buffer is sized from
  // the shape below
  RC(validate_call(s, left_image_pyr, K_pyr, T_right_in_lefts, right_image_l0, right_image_l4));
  B200MVS_CUDA_OK(cudaSetDevice(net->device));
  const Levels L = levels_of(s);
  const size_t B = s.batch, V = s.views, D = s.num_idepth_samples;
  // Staging buffers for one call.
  size_t in_floats = 0;
  for (int l = 0; l < 5; ++l) in_floats += B * 3 * L.px[l] + B * 16;
  in_floats += V * (B * 16 + B * 3 * L.px[0] + B * 3 * L.px[4]);
  size_t out_f = 0, out_m = 0;
  for (int l = 0; l < 5; ++l) {
    out_f += 2 * B * L.px[l];
    out_m += B * D * L.px[l];
  }
  const size_t in_bytes = (in_floats * sizeof(float) + 255) & ~(size_t)255;
  const size_t of_bytes = (out_f * sizeof(float) + 255) & ~(size_t)255;
  const size_t need = in_bytes + of_bytes + out_m;
  if (need > net->host_stage_bytes) {
    if (net->host_stage != nullptr) {
      B200MVS_CUDA_OK(cudaDeviceSynchronize());
      B200MVS_CUDA_OK(cudaFree(net->host_stage));
      net->host_stage = nullptr;
      net->host_stage_bytes = 0;
    }
    if (cudaMalloc(reinterpret_cast<void**>(&net->host_stage), need) != cudaSuccess) {
      set_error("b200mvs_forward_host: staging allocation failed");
      return B200MVS_ENOMEM;
    }
    net->host_stage_bytes = need;
  }
  float* din = reinterpret_cast<float*>(net->host_stage);
  float* dof = reinterpret_cast<float*>(net->host_stage + in_bytes);
  uint8_t* dom = reinterpret_cast<uint8_t*>(net->host_stage + in_bytes + of_bytes);
  cudaStream_t stream = net->host_stream;
  cudaStream_t cs = net->copy_stream;
  static const bool hprof = getenv("B200MVS_HOST_PROFILE") != nullptr;
  auto now = []() { return std::chrono::duration<double, std::micro>(std::chrono::steady_clock::now().time_since_epoch()).count(); };
  const double t_entry = hprof ? now() : 0.0;
  int64_t h2d = 0, d2h = 0;
  int rc = 0;
  auto up = [&](const float* src, size_t count, float*& cursor) -> const float* {
    float* dst = cursor;
    cursor += count;
    if (cudaMemcpyAsync(dst, src, count * sizeof(float), cudaMemcpyHostToDevice, cs) != cudaSuccess) rc = -2;
    h2d += (int64_t)(count * sizeof(float));
    return dst;
  };
  // Uploads in the order the path consumes them, each piece followed by an event the compute streams wait on:
  // cameras (packed into one pinned buffer: one copy instead of 5 + V tiny ones), comparison images at level 0
  // (warp -> right feature network -> recurrence: the critical path), reference image at level 0 (left feature
  // network, side stream), then everything the refiners and the 1/16-scale warps need.
  float* cur = din;
  const float* dl[5];
  const float* dk[5];
  const float *dT[kMaxViews], *dr0[kMaxViews], *dr4[kMaxViews];
  {
    const size_t small = (5 + V) * B * 16;
    if (small > net->pinned_small_floats) {
      if (net->pinned_small != nullptr) {
        B200MVS_CUDA_OK(cudaStreamSynchronize(cs));
        cudaFreeHost(net->pinned_small);
        net->pinned_small = nullptr;
      }
      B200MVS_CUDA_OK(cudaMallocHost(reinterpret_cast<void**>(&net->pinned_small), small * sizeof(float)));
      net->pinned_small_floats = small;
    } else {
      // the previous call's copy out of this buffer completed before that call returned (it synchronised)
    }
    float* hp = net->pinned_small;
    for (int l = 0; l < 5; ++l) {
      std::memcpy(hp + (size_t)l * B * 16, K_pyr[l], B * 16 * sizeof(float));
      dk[l] = cur + (size_t)l * B * 16;
    }
    for (size_t v = 0; v < V; ++v) {
      std::memcpy(hp + (5 + v) * B * 16, T_right_in_lefts[v], B * 16 * sizeof(float));
      dT[v] = cur + (5 + v) * B * 16;
    }
    up(hp, small, cur);
    cudaEventRecord(net->ev_up[0], cs);
  }
  for (size_t v = 0; v < V; ++v) dr0[v] = up(right_image_l0[v], B * 3 * L.px[0], cur);
  cudaEventRecord(net->ev_up[1], cs);
  for (size_t v = 0; v < V; ++v) dr4[v] = up(right_image_l4[v], B * 3 * L.px[4], cur);
  cudaEventRecord(net->ev_up[2], cs);
  dl[0] = up(left_image_pyr[0], B * 3 * L.px[0], cur);
  cudaEventRecord(net->ev_up[3], cs);
  for (int l = 1; l < 5; ++l) dl[l] = up(left_image_pyr[l], B * 3 * L.px[l], cur);
  cudaEventRecord(net->ev_up[4], cs);
  float *oi[5], *orw[5];
  uint8_t* om[5];
  {
    float* c = dof;
    uint8_t* m = dom;
    for (int l = 0; l < 5; ++l) {
      oi[l] = c;
      c += B * L.px[l];
      orw[l] = c;
      c += B * L.px[l];
      // mask volumes: all five when the caller passes no mask list (the reference always computes them); with a
      // list, only the levels it asks for (and level 4, which the view reduction produces anyway)
      om[l] = (out_mask == nullptr || out_mask[l] != nullptr || l == 4) ? m : nullptr;
      m += B * D * L.px[l];
    }
  }
  const double t_up = hprof ? now() : 0.0;
  g_launches = 0;
  if (rc == 0)
    rc = forward_impl(net, net->lanes[0], false, s, dl, dk, dT, dr0, dr4, oi, orw, om, stream, net->ev_up, net->ev_coarse);
  if (rc == 0) {
    net->last_shape = s;
    net->have_last = true;
    net->last_launches = g_launches;
    net->last_lanes = 1;
  }
  const double t_enq = hprof ? now() : 0.0;
  if (rc == 0) {
    // levels 1-4 are final before the level-0 refiner starts: their download runs next to it on the copy stream
    if (net->early_d2h) cudaStreamWaitEvent(cs, net->ev_coarse, 0);
    for (int l = 0; l < 5; ++l) {
      cudaStream_t ds = (l == 0 || !net->early_d2h) ? stream : cs;
      if (out_idepth != nullptr && out_idepth[l] != nullptr) {
        cudaMemcpyAsync(out_idepth[l], oi[l], B * L.px[l] * sizeof(float), cudaMemcpyDeviceToHost, ds);
        d2h += (int64_t)(B * L.px[l] * sizeof(float));
      }
      if (out_idepth_raw != nullptr && out_idepth_raw[l] != nullptr) {
        cudaMemcpyAsync(out_idepth_raw[l], orw[l], B * L.px[l] * sizeof(float), cudaMemcpyDeviceToHost, ds);
        d2h += (int64_t)(B * L.px[l] * sizeof(float));
      }
      if (out_mask != nullptr && out_mask[l] != nullptr) {
        cudaMemcpyAsync(out_mask[l], om[l], B * D * L.px[l], cudaMemcpyDeviceToHost, stream);
        d2h += (int64_t)(B * D * L.px[l]);
      }
    }
    cudaError_t e = cudaStreamSynchronize(cs);
    if (e == cudaSuccess) e = cudaStreamSynchronize(stream);
    if (hprof)
      fprintf(stderr, "forward_host: uploads enqueued %.0f us | kernels enqueued %.0f us | wait for completion %.0f us\n",
              t_up - t_entry, t_enq - t_up, now() - t_enq);
    if (e != cudaSuccess) {
      set_error(std::string("b200mvs_forward_host: ") + cudaGetErrorString(e));
      rc = B200MVS_ECUDA;
    } else {
      rc = check_sweep_watchdog(net);
    }
  } else {
    cudaStreamSynchronize(cs);
    cudaStreamSynchronize(stream);
    if (net->cur->side != nullptr) cudaStreamSynchronize(net->cur->side);   // a failed forward may leave it forked
    if (rc == -2 && g_error.empty()) set_error("b200mvs_forward_host: upload failed");
  }
  if (h2d_bytes != nullptr) *h2d_bytes = h2d;
  if (d2h_bytes != nullptr) *d2h_bytes = d2h;
  return rc;
}

B200MVS_API int b200mvs_get_stage(b200mvs_net* net, const char* name, void* dst, int64_t capacity, int64_t* nbytes,
                      void* stream) {
  if (net == nullptr || name == nullptr || !net->have_last) {
    set_error("b200mvs_get_stage: no forward has run on this handle");
    return B200MVS_EINVAL;
  }
  const b200mvs_shape& s = net->last_shape;
  const Levels L = levels_of(s);
  if (net->last_lanes != 1) {
    set_error("b200mvs_get_stage: the last forward was split into lanes; set option \"lanes\" to 1 to inspect stages");
    return B200MVS_EINVAL;
  }
  const Workspace& ws = net->lanes[0].ws;
  const size_t B = s.batch, V = s.views, D = s.num_idepth_samples, n = B * V;
  const std::string k(name);
  const void* src = nullptr;
  size_t bytes = 0;
  if (k == "idepth_samples") { src = ws.geo.samples; bytes = n * D * 4; }
  else if (k == "baseline") { src = ws.geo.baseline; bytes = n * 4; }
  else if (k == "H0") { src = ws.geo.H0; bytes = n * 9 * 4; }
  else if (k == "H") { src = ws.geo.H; bytes = n * D * 9 * 4; }
  else if (k == "H_inc") { src = ws.geo.Hinc; bytes = n * D * 9 * 4; }
  else if (k == "right_image0_warped") { src = ws.warped0; bytes = n * 3 * L.px[0] * 4; }
  else if (k == "l0_mask") { src = ws.l0mask; bytes = n * L.px[0]; }
  else if (k == "l4_mask") { src = ws.mask_views; bytes = n * D * L.px[4]; }
  else if (k == "left_feature1") { src = ws.f1; bytes = B * L.px[1] * kC * 4; }
  else if (k == "left_feature2") { src = ws.f2; bytes = B * L.px[2] * kC * 4; }
  else if (k == "left_feature3") { src = ws.f3; bytes = B * L.px[3] * kC * 4; }
  else if (k == "left_feature4") { src = ws.feat4; bytes = B * L.px[4] * kC * 4; }
  else if (k == "right_feature_volume") {
    if (!net->keep_stages) {
      set_error("b200mvs_get_stage: right_feature_volume needs b200mvs_set_debug(net, 1) before the forward");
      return B200MVS_EINVAL;
    }
    src = ws.vol; bytes = n * D * L.px[4] * kC * 4;
  }
  else if (k == "cost_filtered") { src = ws.cost1; bytes = n * D * L.px[4] * 4; }
  else if (k == "idepth4_raw_views") { src = ws.raw_views; bytes = n * L.px[4] * 4; }
  else if (k == "recurrence_profile" && net->rec_prof != nullptr) { src = net->rec_prof; bytes = 16 * 12 * 8; }
  else if (k == "recurrence_flags" && ws.rec_flags != nullptr) { src = ws.rec_flags; bytes = n * 17 * 4; }
  else if (k == "recurrence_trace" && net->rec_prof != nullptr) { src = net->rec_prof + 16 * 12; bytes = 16 * 16 * 32 * 8; }
  else {
    set_error("b200mvs_get_stage: unknown stage '" + k + "'");
    return B200MVS_EINVAL;
  }
  if (nbytes != nullptr) *nbytes = (int64_t)bytes;
  if (dst == nullptr) return 0;  // size query
  if ((int64_t)bytes > capacity) {
    set_error("b200mvs_get_stage: destination too small");
    return B200MVS_EINVAL;
  }
  B200MVS_CUDA_OK(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToDevice, static_cast<cudaStream_t>(stream)));
  return 0;
}

B200MVS_API int b200mvs_homography_warp(const float* H, const float* image, int32_t n, int32_t channels, int32_t rows,
                            int32_t cols, int32_t zero_invalid, float* pred, uint8_t* mask, void* stream) {
  if (H == nullptr || image == nullptr || pred == nullptr || n < 0 || channels < 1 || rows < 1 || cols < 1) {
    set_error("b200mvs_homography_warp: bad argument");
    return B200MVS_EINVAL;
  }
  if (n > 65535) {
    set_error("b200mvs_homography_warp: at most 65535 images per call");
    return B200MVS_EINVAL;
  }
  ViewPtrs src{};
  src.views = 1;
  src.p[0] = image;
  return launch_warp_planar(H, 9, src, n, channels, rows, cols, zero_invalid != 0, pred, mask,
                            static_cast<cudaStream_t>(stream));
}

B200MVS_API int b200mvs_upsample_mask(const uint8_t* mask, int64_t planes, int32_t rows, int32_t cols, int32_t out_rows,
                                      int32_t out_cols, int32_t packed, uint8_t* out, void* stream) {
  if (mask == nullptr || out == nullptr || planes < 0 || rows < 1 || cols < 1 || out_rows < 1 || out_cols < 1) {
    set_error("b200mvs_upsample_mask: bad argument");
    return B200MVS_EINVAL;
  }
  return launch_upsample_mask(mask, planes, rows, cols, out_rows, out_cols, out, static_cast<cudaStream_t>(stream),
                              packed != 0);
}

B200MVS_API int b200mvs_reproject(const float* K, const float* T_right_in_left, const float* map, int32_t map_kind,
                                  const float* right_image, int32_t n, int32_t channels, int32_t rows, int32_t cols,
                                  float* pred, uint8_t* mask, float* right_pixels, float* right_idepths,
                                  float* idepth_out, float* disparity_out, void* stream) {
  if (K == nullptr || T_right_in_left == nullptr || map == nullptr || n < 0 || rows < 1 || cols < 1 || map_kind < 0 ||
      map_kind > 2 || (pred != nullptr && (right_image == nullptr || channels < 1))) {
    set_error("b200mvs_reproject: bad argument");
    return B200MVS_EINVAL;
  }
  if (n > 65535) {
    set_error("b200mvs_reproject: at most 65535 images per call");
    return B200MVS_EINVAL;
  }
  return launch_reproject(K, T_right_in_left, map, map_kind, right_image, n, channels, rows, cols, pred, mask,
                          right_pixels, right_idepths, idepth_out, disparity_out, static_cast<cudaStream_t>(stream));
}

B200MVS_API int b200mvs_area_downsample(const float* in, int32_t planes, int32_t rows, int32_t cols, float* out,
                                        void* stream) {
  if (in == nullptr || out == nullptr || planes < 0 || rows < 1 || cols < 1) {
    set_error("b200mvs_area_downsample: bad argument");
    return B200MVS_EINVAL;
  }
  return launch_area_downsample(in, planes, rows, cols, out, static_cast<cudaStream_t>(stream));
}

B200MVS_API int b200mvs_prepare_cameras(const float* K, const float* const* T_right_in_lefts, int32_t batch,
                                        int32_t views, int32_t levels, const int32_t* level_sizes, float* K_pyr,
                                        float* T_right_in_left_out, float* T_left_in_right_out, float* baseline,
                                        void* stream_) {
  if (K == nullptr || T_right_in_lefts == nullptr || level_sizes == nullptr || K_pyr == nullptr ||
      T_right_in_left_out == nullptr || T_left_in_right_out == nullptr || baseline == nullptr || batch < 0 ||
      views < 1 || views > kMaxViews || levels < 1 || levels > 16) {
    set_error("b200mvs_prepare_cameras: bad argument");
    return B200MVS_EINVAL;
  }
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  ViewPtrs T{};
  T.views = views;
  for (int v = 0; v < views; ++v) {
    if (T_right_in_lefts[v] == nullptr) {
      set_error("b200mvs_prepare_cameras: missing pose");
      return B200MVS_EINVAL;
    }
    T.p[v] = T_right_in_lefts[v];
  }
  int* dsizes = nullptr;
  B200MVS_CUDA_OK(cudaMallocAsync(reinterpret_cast<void**>(&dsizes), sizeof(int) * 2 * levels, stream));
  B200MVS_CUDA_OK(cudaMemcpyAsync(dsizes, level_sizes, sizeof(int) * 2 * levels, cudaMemcpyHostToDevice, stream));
  const int rc = launch_prepare_cameras(K, T, batch, levels, dsizes, K_pyr, T_right_in_left_out, T_left_in_right_out,
                                        baseline, stream);
  cudaFreeAsync(dsizes, stream);
  return rc;
}

B200MVS_API int b200mvs_depth_metrics(const float* est, const float* baseline, const float* depth_true,
                                      int32_t est_is_depth, float min_depth, float max_depth, int32_t batch,
                                      int64_t pixels, float* idepth_out, float* depth_out, double* metrics,
                                      void* stream) {
  if (est == nullptr || batch < 1 || pixels < 1) {
    set_error("b200mvs_depth_metrics: bad argument");
    return B200MVS_EINVAL;
  }
  if (metrics != nullptr && depth_true == nullptr) {
    set_error("b200mvs_depth_metrics: metrics need depth_true");
    return B200MVS_EINVAL;
  }
  return launch_depth_metrics(est, baseline, depth_true, est_is_depth != 0, min_depth, max_depth, batch, pixels,
                              idepth_out, depth_out, metrics, static_cast<cudaStream_t>(stream));
}

}  // extern "C"
