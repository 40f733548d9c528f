// Network orchestrator of libboa_b200: builds the static launch schedule of one PlainConvUNet forward over a batch
// of sliding-window patches and runs it (plain launches or a captured CUDA graph).
//
// Replaces  nnUNetPredictor.initialize_from_trained_model_folder  (_external/nnunetv2/inference/predict_from_raw_data.py:67-129)
//           get_network_from_plans -> PlainConvUNet(**kwargs)     (_external/nnunetv2/utilities/get_network_from_plans.py:9-43,
//                                                                  _external/nnunetv2/utilities/plans_handling/plans_handler.py:36-97)
//           `self.network(x)` and the per-patch Gaussian accumulation (predict_from_raw_data.py:543,603-616)
// Network semantics follow dynamic_network_architectures==0.4.3 (PlainConvEncoder / UNetDecoder / StackedConvBlocks /
// ConvDropoutNormReLU); see oracle/network.py for the restatement this is tested against.
//
// Numerics: fp16 operands, fp32 accumulation on the tensor cores; conv outputs are stored once as fp16; InstanceNorm
// statistics come from the fp32 accumulators (fp64 sums); normalise + LeakyReLU are applied to the stored fp16 values
// in fp32 and stored as fp16 (the reference's CUDA path runs the same ops under torch.autocast fp16,
// predict_from_raw_data.py:648).
#include <string.h>
#include <algorithm>
#include <map>
#include <string>
#include <vector>
#include "net_kernels.cuh"

namespace boa {

struct HostTensor {
  std::vector<float> data;
  std::vector<int64_t> shape;
};

enum StepKind { STEP_CONV_FOLD, STEP_CONV_TAPS, STEP_CONV_SIMT, STEP_TCONV_TAPS, STEP_TCONV_SIMT, STEP_CONV_FIRST };

struct ConvStep {  // conv -> statistics -> (unfused schedule only: norm + lrelu pass)   |   transposed conv
  StepKind kind;
  bool is_tconv = false;
  ActView src;           // input activation (for CONV_TAPS stride 2: the s2d view)
  ActView src_plain;     // SIMT stride-2 path reads the plain activation
  InXform xf;            // normalisation of the INPUT applied by this step's kernel (fused schedule; scale == nullptr: none)
  int cin = 0, cout = 0;
  int ks[3] = {3, 3, 3}, stride[3] = {1, 1, 1};
  int Do = 0, Ho = 0, Wo = 0;
  ActView out;                 // where the kernel writes: RAW conv output / transposed-conv output
  __half* out_s2d = nullptr;   // fused schedule: space-to-depth copy of the raw output, written by the conv epilogue
  bool norm_pass = false;      // unfused schedule: a standalone normalise pass follows (out -> dst [+ s2d])
  ActView dst;                 // unfused schedule: normalised output
  __half* s2d = nullptr;       // unfused schedule: space-to-depth copy of the normalised output
  float *d_w = nullptr, *d_bias = nullptr, *d_gamma = nullptr, *d_beta = nullptr;  // SIMT operands / norm affine
  double* d_stats = nullptr;
  float *d_scale = nullptr, *d_shift = nullptr;
  float *d_scale2 = nullptr, *d_shift2 = nullptr;  // second copy of scale / shift: rows of a concat's combined table
  int stride2 = 0, off2 = 0;
  ConvMmaPlan* fold = nullptr;
  ConvTapsPlan* taps = nullptr;
  double macs = 0;
  std::string name;
};

// Scratch memory of one forward (activations, statistics, staging).  Networks of identical geometry run one at a
// time on a stream and may share it (fold ensembles, the five `total` part models).
// Two LANES: consecutive batches of patches alternate between two complete sets of scratch buffers on two internal
// streams, so the HBM-bound passes of one batch overlap the tensor-bound convolutions of the other and kernel
// boundaries stop draining the machine; only the head (the ordered `logits[sl] += pred * g`) is serialised between
// lanes with an event.  The streams belong to the workspace, so networks that share it stay ordered.
constexpr int MAX_LANES = 4;
struct Workspace {
  std::vector<std::pair<void*, size_t>> bufs[MAX_LANES];
  cudaStream_t stream[MAX_LANES] = {};
  cudaEvent_t head_done = nullptr;   // last head enqueued on any lane
  cudaEvent_t fork = nullptr, join[MAX_LANES] = {};
  int last_lane = -1;
  int refs = 1;
};

struct Lane {
  std::vector<ConvStep> steps;
  __half* d_patch = nullptr;  // [B][2][P] C8 input (or plain fp16 [B][P], see input_mode)
  FwdCall* d_call = nullptr;
  double* d_stats_all = nullptr;
  ActView head_src, head_src_raw;  // normalised / raw output of the last conv (fused head normalisation)
  const float *head_scale = nullptr, *head_shift = nullptr;
  cudaGraphExec_t g_body = nullptr, g_head = nullptr;
  int launches_body = 0, launches_head = 0;
  size_t ws_cursor = 0;
};

}  // namespace boa

using namespace boa;

struct boa_net {
  boa_arch arch;
  int device = 0, B = 1, mode = 0;
  bool finalized = false;
  std::map<std::string, HostTensor> tensors;
  std::vector<void*> allocs;  // per-network memory (weights)
  Workspace* ws = nullptr;    // shared scratch
  Lane lane[MAX_LANES];
  int n_lanes = 2;  // BOA_B200_LANES (1..MAX_LANES)
  bool fuse = true;  // fused schedule: every kernel normalises its input itself, no standalone normalise pass
                     // (BOA_B200_UNFUSED=1 at creation selects the pass-per-layer schedule: cross-check / A-B timing)
  std::map<std::string, float*> wcache;  // device copies of the fp32 parameters, shared by the lanes
  // head
  float *d_head_w = nullptr, *d_head_b = nullptr;
  int input_mode = 0;  // 0: C8 16ch, 1: plain fp16 (direct first layer), 2: C8 with the 9 in-plane neighbours on K
  static constexpr int CALL_RING = 8;
  FwdCall* h_call = nullptr;  // pinned ring of staging slots, one per batch in flight
  cudaEvent_t call_ev[CALL_RING] = {};
  int call_slot = 0;
  size_t stats_bytes = 0;
  int64_t macs_per_patch = 0;
  // timing
  bool timing = false;
  std::vector<cudaEvent_t> ev;
  size_t ev_used = 0;
  std::vector<std::pair<size_t, size_t>> conv_spans;  // event index pairs around conv kernels
  std::vector<std::pair<int, double>> conv_span_info;  // (StepKind, algorithmic FLOP of that launch) per span
  std::vector<std::pair<size_t, size_t>> fwd_spans;
  double ms_convs = 0, ms_total = 0;
  int64_t n_conv_launches = 0;
  int use_graph = 1;
  int cur_nb = 0;      // batch items the next launches process (plain launches: the real count; graphs: always B)
  int debug_only = 0;  // profiling aid (BOA_B200_DEBUG_ONLY=conv|thin at creation): launch only one class of kernels
};

namespace {

template <typename T>
T* dalloc(boa_net* net, size_t n) {
  void* p = nullptr;
  if (cudaMalloc(&p, n * sizeof(T)) != cudaSuccess) {
    set_error("cudaMalloc of %zu bytes failed: %s", n * sizeof(T), cudaGetErrorString(cudaGetLastError()));
    return nullptr;
  }
  net->allocs.push_back(p);
  return static_cast<T*>(p);
}

// Workspace allocation: buffers are requested in a deterministic order, so a network that shares a donor's workspace
// walks the donor's list and must find the same sizes.
template <typename T>
T* wsalloc(boa_net* net, int li, size_t n) {
  const size_t bytes = n * sizeof(T);
  Workspace* ws = net->ws;
  size_t& cursor = net->lane[li].ws_cursor;
  std::vector<std::pair<void*, size_t>>& bufs = ws->bufs[li];
  if (cursor < bufs.size()) {
    auto& b = bufs[cursor];
    if (b.second != bytes) {
      set_error("shared workspace mismatch at buffer %zu: have %zu bytes, need %zu (different network geometry)",
                cursor, b.second, bytes);
      return nullptr;
    }
    ++cursor;
    return static_cast<T*>(b.first);
  }
  void* p = nullptr;
  if (cudaMalloc(&p, bytes) != cudaSuccess) {
    set_error("cudaMalloc of %zu bytes failed: %s", bytes, cudaGetErrorString(cudaGetLastError()));
    return nullptr;
  }
  bufs.push_back({p, bytes});
  ++cursor;
  return static_cast<T*>(p);
}

// fp32 parameters are uploaded once and shared by the lanes
float* upload(boa_net* net, const std::string& key, const std::vector<float>& v) {
  auto it = net->wcache.find(key);
  if (it != net->wcache.end()) return it->second;
  float* d = dalloc<float>(net, v.size());
  if (d) cudaMemcpy(d, v.data(), v.size() * sizeof(float), cudaMemcpyHostToDevice);
  net->wcache[key] = d;
  return d;
}

std::vector<float> round_fp16(const std::vector<float>& v) {
  std::vector<float> o(v.size());
  for (size_t i = 0; i < v.size(); ++i) o[i] = __half2float(__float2half_rn(v[i]));
  return o;
}

const HostTensor* find(boa_net* net, const std::string& key) {
  auto it = net->tensors.find(key);
  return it == net->tensors.end() ? nullptr : &it->second;
}

bool is3(const int* a, int v) { return a[0] == v && a[1] == v && a[2] == v; }

cudaEvent_t next_event(boa_net* net, size_t* idx) {
  if (net->ev_used == net->ev.size()) {
    cudaEvent_t e;
    cudaEventCreate(&e);
    net->ev.push_back(e);
  }
  *idx = net->ev_used;
  return net->ev[net->ev_used++];
}

int run_step(boa_net* net, ConvStep& st, cudaStream_t s) {
  const boa_arch& a = net->arch;
  const int B = net->cur_nb > 0 ? net->cur_nb : net->B;
  size_t e0 = 0, e1 = 0;
  const bool time_it = net->timing;
  const bool pass_only = net->debug_only == 2;  // profiling aid: the HBM-bound passes without the convolutions
  if (time_it) cudaEventRecord(next_event(net, &e0), s);
  int r = BOA_OK;
  if (st.is_tconv) {
    if (pass_only) return BOA_OK;
    if (st.kind == STEP_TCONV_TAPS && net->mode == 0) r = conv_taps_launch(st.taps, s, B);
    else r = launch_tconv_simt(st.src, B, st.d_w, st.d_bias, st.cin, st.cout, st.stride, st.out, s, st.xf);
    if (time_it) {
      cudaEventRecord(next_event(net, &e1), s);
      net->conv_spans.push_back({e0, e1});
      net->conv_span_info.push_back({net->mode == 0 ? (int)st.kind : (int)STEP_TCONV_SIMT, 2.0 * st.macs * B});
    }
    return r;
  }
  if (!pass_only) {
    if (st.kind == STEP_CONV_FIRST)
      r = launch_conv_first(st.src.base, B, st.d_w, st.d_bias, st.cout, st.out.base, st.Do, st.Ho, st.Wo, st.d_stats, s);
    else if (net->mode == 0 && st.kind == STEP_CONV_FOLD) r = conv_mma_launch(st.fold, s, B);
    else if (net->mode == 0 && st.kind == STEP_CONV_TAPS) r = conv_taps_launch(st.taps, s, B);
    else
      r = launch_conv_simt(st.src_plain, B, st.d_w, st.d_bias, st.cin, st.cout, st.ks, st.stride, st.out, st.Do, st.Ho,
                           st.Wo, st.d_stats, s, st.xf);
    if (time_it) {
      cudaEventRecord(next_event(net, &e1), s);
      net->conv_spans.push_back({e0, e1});
      net->conv_span_info.push_back({(net->mode == 0 || st.kind == STEP_CONV_FIRST) ? (int)st.kind : (int)STEP_CONV_SIMT,
                                     2.0 * st.macs * B});
    }
    if (r) return r;
    if (net->debug_only == 1) return BOA_OK;  // convolutions only
  }
  if ((r = launch_stats_finalize(st.d_stats, st.d_gamma, st.d_beta, B, st.cout, (double)st.Do * st.Ho * st.Wo, a.eps,
                                 st.d_scale, st.d_shift, s, st.d_scale2, st.d_shift2, st.stride2, st.off2)))
    return r;
  if (!st.norm_pass) return BOA_OK;
  return launch_norm_lrelu(st.out.base, B, st.cout / 8, st.Do, st.Ho, st.Wo, st.d_scale, st.d_shift, a.leaky_slope,
                           st.dst, st.s2d, s);
}

// everything of one forward except the input staging and the head
int run_body(boa_net* net, Lane& L, cudaStream_t s) {
  BOA_CUDA(cudaMemsetAsync(L.d_stats_all, 0, net->stats_bytes, s));
  for (ConvStep& st : L.steps)
    if (int r = run_step(net, st, s)) return r;
  return BOA_OK;
}

// input staging + network body of one batch
int run_front(boa_net* net, Lane& L, cudaStream_t s) {
  const boa_arch& a = net->arch;
  if (int r = launch_extract_patches(L.d_call, net->cur_nb > 0 ? net->cur_nb : net->B, a.patch[0], a.patch[1], a.patch[2], L.d_patch, net->input_mode, s))
    return r;
  return run_body(net, L, s);
}

// heads of one batch, one launch per patch in slicer order; d_logits != nullptr: raw logits of nb patches instead
int run_heads(boa_net* net, Lane& L, int nb, float* d_logits, cudaStream_t s) {
  const boa_arch& a = net->arch;
  const bool fh = L.head_scale != nullptr;
  const size_t pv = (size_t)a.patch[0] * a.patch[1] * a.patch[2];
  if (net->debug_only == 1) return BOA_OK;
  const bool time_it = net->timing && !d_logits;
  for (int b = 0; b < nb; ++b) {
    size_t e0 = 0, e1 = 0;
    if (time_it) cudaEventRecord(next_event(net, &e0), s);
    if (int r = launch_head(fh ? L.head_src_raw : L.head_src, b, net->d_head_w, net->d_head_b, a.features[0],
                            a.num_classes, d_logits ? d_logits + (size_t)b * a.num_classes * pv : nullptr, L.d_call,
                            fh ? L.head_scale : nullptr, fh ? L.head_shift : nullptr, a.leaky_slope, s))
      return r;
    if (time_it) {  // kind 6: the second field carries the launch's algorithmic BYTES (activations read, fp32
                    // read-modify-write of the C logit planes, Gaussian map)
      cudaEventRecord(next_event(net, &e1), s);
      net->conv_spans.push_back({e0, e1});
      net->conv_span_info.push_back({6, (double)pv * (2.0 * a.features[0] + 8.0 * a.num_classes + 4.0)});
    }
  }
  return BOA_OK;
}

void destroy_graphs(boa_net* net) {
  for (Lane& L : net->lane) {
    if (L.g_body) cudaGraphExecDestroy(L.g_body);
    if (L.g_head) cudaGraphExecDestroy(L.g_head);
    L.g_body = L.g_head = nullptr;
  }
}

template <typename F>
int capture_graph(cudaGraphExec_t* out, int* launches, F&& body) {
  cudaStream_t cs;
  BOA_CUDA(cudaStreamCreateWithFlags(&cs, cudaStreamNonBlocking));
  cudaGraph_t g = nullptr;
  BOA_CUDA(cudaStreamBeginCapture(cs, cudaStreamCaptureModeThreadLocal));
  const uint64_t before = g_launches.load();
  const int r = body(cs);
  *launches = (int)(g_launches.load() - before);
  const cudaError_t ce = cudaStreamEndCapture(cs, &g);
  cudaStreamDestroy(cs);
  g_launches.fetch_sub((uint64_t)*launches);  // the capture enqueued nothing
  if (r) return r;
  BOA_CUDA(ce);
  BOA_CUDA(cudaGraphInstantiate(out, g, 0));
  cudaGraphDestroy(g);
  return BOA_OK;
}

}  // namespace

extern "C" int boa_net_create(const boa_arch* arch, int device, int max_batch, boa_net** out) {
  BOA_REQUIRE(arch && out, "boa_net_create: null argument");
  BOA_REQUIRE(arch->n_stages >= 2 && arch->n_stages <= BOA_MAX_STAGES, "boa_net_create: n_stages=%d unsupported",
              arch->n_stages);
  BOA_REQUIRE(max_batch >= 1 && max_batch <= MAX_BATCH, "boa_net_create: max_batch must be in [1, %d]", MAX_BATCH);
  BOA_REQUIRE(arch->in_channels == 1, "boa_net_create: only single-channel (CT) input is implemented");
  BOA_REQUIRE(arch->num_classes >= 1 && arch->num_classes <= 128, "boa_net_create: num_classes=%d out of range",
              arch->num_classes);
  for (int s = 0; s < arch->n_stages; ++s) {
    BOA_REQUIRE(arch->features[s] % 8 == 0 && arch->features[s] > 0, "boa_net_create: features must be multiples of 8");
    for (int k = 0; k < 3; ++k) {
      BOA_REQUIRE(arch->kernels[s][k] == 1 || arch->kernels[s][k] == 3, "boa_net_create: kernel sizes must be 1 or 3");
      BOA_REQUIRE(arch->strides[s][k] == 1 || arch->strides[s][k] == 2, "boa_net_create: strides must be 1 or 2");
    }
  }
  BOA_REQUIRE(arch->features[0] <= 64, "boa_net_create: features[0] > 64 is not supported by the head kernel");
  int ndev = 0;
  BOA_CUDA(cudaGetDeviceCount(&ndev));
  BOA_REQUIRE(device >= 0 && device < ndev, "boa_net_create: device %d not present (%d devices)", device, ndev);
  BOA_CUDA(cudaSetDevice(device));
  boa_net* net = new boa_net();
  net->arch = *arch;
  net->device = device;
  net->B = max_batch;
  net->ws = new Workspace();
  const char* nl = getenv("BOA_B200_LANES");
  net->n_lanes = nl ? std::max(1, std::min(MAX_LANES, atoi(nl))) : 2;
  if (const char* u = getenv("BOA_B200_UNFUSED")) net->fuse = atoi(u) == 0;
  if (const char* d = getenv("BOA_B200_DEBUG_ONLY")) net->debug_only = !strcmp(d, "conv") ? 1 : (!strcmp(d, "thin") ? 2 : 0);
  *out = net;
  return BOA_OK;
}

extern "C" int boa_net_set_tensor(boa_net* net, const char* key, const float* h_data, const int64_t* shape, int ndim) {
  BOA_REQUIRE(net && key && h_data && shape, "boa_net_set_tensor: null argument");
  BOA_REQUIRE(!net->finalized, "boa_net_set_tensor: network already finalized");
  const std::string k(key);
  // aliases of the same parameters registered twice by the package (StackedConvBlocks.all_modules, UNetDecoder.encoder)
  if (k.find("all_modules") != std::string::npos || k.rfind("decoder.encoder.", 0) == 0) return BOA_OK;
  HostTensor t;
  size_t n = 1;
  for (int i = 0; i < ndim; ++i) {
    t.shape.push_back(shape[i]);
    n *= (size_t)shape[i];
  }
  t.data.assign(h_data, h_data + n);
  net->tensors[k] = std::move(t);
  return BOA_OK;
}

extern "C" int boa_net_share_workspace(boa_net* net, boa_net* donor) {
  BOA_REQUIRE(net && donor && net != donor, "boa_net_share_workspace: bad argument");
  BOA_REQUIRE(!net->finalized && net->ws->bufs[0].empty(), "boa_net_share_workspace: must be called before finalize");
  BOA_REQUIRE(net->device == donor->device && net->B == donor->B, "boa_net_share_workspace: device / batch differ");
  if (--net->ws->refs == 0) delete net->ws;
  net->ws = donor->ws;
  ++net->ws->refs;
  return BOA_OK;
}

extern "C" int boa_net_set_mode(boa_net* net, int mode) {
  BOA_REQUIRE(net && (mode == 0 || mode == 1), "boa_net_set_mode: bad argument");
  if (net->mode != mode) destroy_graphs(net);
  net->mode = mode;
  return BOA_OK;
}

extern "C" int boa_net_set_graph(boa_net* net, int enable) {
  BOA_REQUIRE(net, "boa_net_set_graph: null");
  net->use_graph = enable ? 1 : 0;
  return BOA_OK;
}

static int need(boa_net* net, const std::string& key, const HostTensor** out, size_t expect) {
  const HostTensor* t = find(net, key);
  if (!t) {
    set_error("boa_net_finalize: missing tensor '%s'", key.c_str());
    return BOA_ERR_STATE;
  }
  if (t->data.size() != expect) {
    set_error("boa_net_finalize: tensor '%s' has %zu elements, expected %zu", key.c_str(), t->data.size(), expect);
    return BOA_ERR_ARG;
  }
  *out = t;
  return BOA_OK;
}

// What a consumer reads: a C8 tensor plus, in the fused schedule, the normalisation still to be applied to it.
struct Feed {
  ActView view;
  InXform xf;
};

static int build_lane(boa_net* net, int li) {
  Lane& L = net->lane[li];
  const boa_arch& a = net->arch;
  const int n = a.n_stages, B = net->B;
  const bool fuse = net->fuse;
  // ---- spatial dims per stage
  int dims[BOA_MAX_STAGES][3];
  for (int k = 0; k < 3; ++k) dims[0][k] = a.patch[k];
  for (int s = 0; s < n; ++s) {
    for (int k = 0; k < 3; ++k) {
      const int prev = s == 0 ? a.patch[k] : dims[s - 1][k];
      BOA_REQUIRE(prev % a.strides[s][k] == 0, "boa_net_finalize: patch size not divisible by the strides");
      dims[s][k] = prev / a.strides[s][k];
    }
  }
  BOA_REQUIRE(is3(a.strides[0], 1), "boa_net_finalize: stage 0 must have stride 1");
  auto vox = [&](int s) { return (size_t)dims[s][0] * dims[s][1] * dims[s][2]; };
  auto view = [&](__half* base, int groups_total, int off, int groups, int s) {
    ActView v;
    v.base = base; v.groups_total = groups_total; v.group_off = off; v.groups = groups;
    v.D = dims[s][0]; v.H = dims[s][1]; v.W = dims[s][2];
    return v;
  };
  // ---- buffers.  Fused schedule: only RAW tensors exist - two ping-pong buffers per stage, the concat buffers (the
  // transposed conv writes the lower half, the encoder's last conv of the stage writes its raw output straight into
  // the upper half) and the space-to-depth copies.  Unfused schedule: plus a normalised copy of everything.
  L.d_patch = wsalloc<__half>(net, li, (size_t)B * 16 * vox(0));
  if (!L.d_patch) return BOA_ERR_CUDA;
  __half *raw[BOA_MAX_STAGES][2], *mid[BOA_MAX_STAGES][2] = {}, *outb[BOA_MAX_STAGES] = {}, *cat[BOA_MAX_STAGES],
      *s2d[BOA_MAX_STAGES];
  float *comb_scale[BOA_MAX_STAGES] = {}, *comb_shift[BOA_MAX_STAGES] = {};
  for (int s = 0; s < n; ++s) {
    const size_t f = (size_t)a.features[s];
    raw[s][0] = wsalloc<__half>(net, li, B * f * vox(s));
    raw[s][1] = wsalloc<__half>(net, li, B * f * vox(s));
    if (!raw[s][0] || !raw[s][1]) return BOA_ERR_CUDA;
    if (!fuse) {
      mid[s][0] = wsalloc<__half>(net, li, B * f * vox(s));
      mid[s][1] = wsalloc<__half>(net, li, B * f * vox(s));
      outb[s] = wsalloc<__half>(net, li, B * f * vox(s));
      if (!mid[s][0] || !mid[s][1] || !outb[s]) return BOA_ERR_CUDA;
    }
    cat[s] = s < n - 1 ? wsalloc<__half>(net, li, B * 2 * f * vox(s)) : nullptr;
    // space-to-depth copy for the next stage's strided conv: any mix of strides 1 / 2 with equal in-plane strides
    // (the unfused schedule's normalise pass only writes the isotropic stride-2 copy)
    bool want_s2d = false;
    if (s < n - 1) {
      const int* ns = a.strides[s + 1];
      want_s2d = !is3(ns, 1) && ns[1] == ns[2] && (fuse || is3(ns, 2));
      for (int k = 0; k < 3; ++k) want_s2d = want_s2d && (ns[k] == 1 || ns[k] == 2) && dims[s][k] % ns[k] == 0;
    }
    s2d[s] = want_s2d ? wsalloc<__half>(net, li, B * f * vox(s)) : nullptr;
    if ((s < n - 1 && !cat[s]) || (want_s2d && !s2d[s])) return BOA_ERR_CUDA;
    if (fuse && s < n - 1) {
      // combined scale / shift rows of the decoder concat [B][2f]: the transposed-conv half is final (identity, never
      // read: ident_groups), the skip half is written by the encoder conv's statistics kernel every forward
      comb_scale[s] = wsalloc<float>(net, li, (size_t)B * 2 * f);
      comb_shift[s] = wsalloc<float>(net, li, (size_t)B * 2 * f);
      if (!comb_scale[s] || !comb_shift[s]) return BOA_ERR_CUDA;
      BOA_CUDA(cudaMemset(comb_scale[s], 0, (size_t)B * 2 * f * sizeof(float)));
      BOA_CUDA(cudaMemset(comb_shift[s], 0, (size_t)B * 2 * f * sizeof(float)));
    }
  }
  // ---- schedule
  int n_norm_layers = 0;
  for (int s = 0; s < n; ++s) n_norm_layers += a.n_conv_enc[s];
  for (int j = 0; j < n - 1; ++j) n_norm_layers += a.n_conv_dec[j];
  int cmax = 0;
  for (int s = 0; s < n; ++s) cmax = std::max(cmax, a.features[s]);
  net->stats_bytes = (size_t)n_norm_layers * B * cmax * 2 * sizeof(double);
  L.d_stats_all = wsalloc<double>(net, li, (size_t)n_norm_layers * B * cmax * 2);
  if (!L.d_stats_all) return BOA_ERR_CUDA;
  int layer_idx = 0;
  double macs = 0;

  // first layer (Cin = 1, 3x3x3): on the tensor cores with the 9 in-plane taps moved onto K (K = 16, three folded
  // dz taps) when Cout % 32 == 0, else the direct FP32 kernel from a plain fp16 patch
  const bool first33 = a.in_channels == 1 && a.kernels[0][1] == 3 && a.kernels[0][2] == 3 &&
                       (a.kernels[0][0] == 3 || a.kernels[0][0] == 1) && a.n_conv_enc[0] >= 1;
  // (the direct kernel writes a dense tensor: not usable when the first conv is also the last of its stage in the
  // fused schedule, where the output goes into the concat buffer)
  const bool first_dense = !(fuse && a.n_conv_enc[0] == 1 && n > 1);
  net->input_mode = (first33 && a.features[0] % 32 == 0) ? 2
                    : (first33 && is3(a.kernels[0], 3) && first_dense && conv_first_supported(a.features[0]) ? 1 : 0);
  bool first_plain = net->input_mode == 1;
  bool first_nb9 = net->input_mode == 2;
  int raw_flip = 0;
  // final_dst: where the consumer expects this conv's result (fused: the raw output is written there; unfused: the
  // normalise pass writes there and the raw output goes to a ping-pong buffer).  want_s2d: the next stage's stride-2
  // conv wants a space-to-depth copy of the result.
  auto add_conv = [&](const std::string& prefix, const Feed& in, const ActView& in_s2d, int cin, int cout,
                      const int* ks, const int* stride, int s_out, const ActView& final_dst, __half* want_s2d,
                      const int* next_stride) -> int {
    ConvStep st;
    st.name = prefix;
    st.cin = cin; st.cout = cout;
    for (int k = 0; k < 3; ++k) { st.ks[k] = ks[k]; st.stride[k] = stride[k]; }
    st.Do = dims[s_out][0]; st.Ho = dims[s_out][1]; st.Wo = dims[s_out][2];
    if (fuse) {
      st.out = final_dst;
    } else {
      st.out = view(raw[s_out][raw_flip & 1], cout / 8, 0, cout / 8, s_out);
      ++raw_flip;
      st.norm_pass = true;
      st.dst = final_dst;
      st.s2d = want_s2d;
    }
    st.src_plain = in.view;
    st.xf = in.xf;
    const size_t k3 = (size_t)ks[0] * ks[1] * ks[2];
    const HostTensor *w, *bi, *g, *be;
    if (int r = need(net, prefix + ".conv.weight", &w, (size_t)cout * cin * k3)) return r;
    if (int r = need(net, prefix + ".conv.bias", &bi, cout)) return r;
    if (int r = need(net, prefix + ".norm.weight", &g, cout)) return r;
    if (int r = need(net, prefix + ".norm.bias", &be, cout)) return r;
    const std::vector<float> wr = round_fp16(w->data);
    st.d_w = upload(net, prefix + ".conv.weight", wr);
    st.d_bias = upload(net, prefix + ".conv.bias", bi->data);
    st.d_gamma = upload(net, prefix + ".norm.weight", g->data);
    st.d_beta = upload(net, prefix + ".norm.bias", be->data);
    st.d_stats = L.d_stats_all + (size_t)layer_idx * B * cmax * 2;
    st.d_scale = wsalloc<float>(net, li, (size_t)B * cout);
    st.d_shift = wsalloc<float>(net, li, (size_t)B * cout);
    if (!st.d_w || !st.d_bias || !st.d_gamma || !st.d_beta || !st.d_scale || !st.d_shift) return BOA_ERR_CUDA;
    ++layer_idx;
    st.macs = (double)k3 * cin * cout * st.Do * st.Ho * st.Wo;
    macs += st.macs;
    st.kind = STEP_CONV_SIMT;
    st.src = in.view;
    ConvIO io;
    io.out = st.out;
    io.xf = in.xf;
    if (next_stride)
      for (int k = 0; k < 3; ++k) io.s2d_stride[k] = next_stride[k];
    const int cin_padded = (cin + 15) / 16 * 16;
    bool k13 = ks[1] == ks[2] && stride[1] == stride[2];
    for (int k = 0; k < 3; ++k) k13 = k13 && (ks[k] == 1 || ks[k] == 3) && (stride[k] == 1 || stride[k] == 2);
    // a transform in the tensor-core kernels works on whole K chunks of 16 channels
    const bool xf_ok = !in.xf.scale || cin % 16 == 0;
    if (first_plain) {
      st.kind = STEP_CONV_FIRST;
    } else if (first_nb9) {
      io.s2d = fuse ? want_s2d : nullptr;
      st.fold = conv_mma_plan_create(wr.data(), bi->data.data(), cin, cin_padded, cout, in.view, B, io, st.d_stats, true,
                                     ks[0]);
      if (!st.fold) return BOA_ERR_CUDA;
      st.kind = STEP_CONV_FOLD;
    } else if (is3(ks, 3) && is3(stride, 1) && cout % 32 == 0 && cin_padded <= in.view.groups * 8 && xf_ok) {
      io.s2d = fuse ? want_s2d : nullptr;
      st.fold = conv_mma_plan_create(wr.data(), bi->data.data(), cin, cin_padded, cout, in.view, B, io, st.d_stats);
      if (!st.fold) return BOA_ERR_CUDA;
      st.kind = STEP_CONV_FOLD;
    } else if (k13 && cin % 16 == 0 && cout % 32 == 0 && xf_ok && (is3(stride, 1) || in_s2d.base)) {
      // every other kernel / stride mix of {1,3} x {1,2} (equal in-plane): the tap-list kernel, on the plain tensor
      // when all strides are 1, else on the space-to-depth copy the producer's epilogue (or pass) wrote
      const ActView& tsrc = is3(stride, 1) ? in.view : in_s2d;
      io.s2d = fuse ? want_s2d : nullptr;
      TapsGeom geo;
      for (int k = 0; k < 3; ++k) { geo.ks[k] = ks[k]; geo.stride[k] = stride[k]; }
      st.taps = conv_taps_plan_create(TAPS_CONV, geo, wr.data(), bi->data.data(), cin, cout, tsrc, B, io, st.d_stats);
      if (!st.taps) return BOA_ERR_CUDA;
      st.kind = STEP_CONV_TAPS;
      st.src = tsrc;
    }
    if (fuse && (st.kind == STEP_CONV_FOLD || st.kind == STEP_CONV_TAPS)) st.out_s2d = want_s2d;
    first_plain = false;
    first_nb9 = false;
    L.steps.push_back(st);
    return BOA_OK;
  };
  // what the consumers of the step just added read
  auto feed_of_last = [&](const ActView& final_dst) {
    Feed f;
    f.view = final_dst;
    if (fuse) {
      const ConvStep& st = L.steps.back();
      f.xf.scale = st.d_scale; f.xf.shift = st.d_shift; f.xf.channels = st.cout; f.xf.ident_groups = 0;
      f.xf.slope = a.leaky_slope;
    }
    return f;
  };

  // encoder
  Feed cur;
  cur.view = view(L.d_patch, 2, 0, 2, 0);
  ActView cur_s2d;  // s2d copy of `cur` when it exists
  int cur_c = a.in_channels;
  for (int s = 0; s < n; ++s) {
    const int f = a.features[s];
    for (int i = 0; i < a.n_conv_enc[s]; ++i) {
      const bool last = i == a.n_conv_enc[s] - 1;
      const int one[3] = {1, 1, 1};
      const int* stride = i == 0 ? a.strides[s] : one;
      ActView dst;
      if (last && s < n - 1) dst = view(cat[s], 2 * f / 8, f / 8, f / 8, s);
      else if (fuse) { dst = view(raw[s][raw_flip & 1], f / 8, 0, f / 8, s); ++raw_flip; }
      else dst = last ? view(outb[s], f / 8, 0, f / 8, s) : view(mid[s][i & 1], f / 8, 0, f / 8, s);
      __half* s2d_out = (last && s < n - 1) ? s2d[s] : nullptr;
      char name[64];
      snprintf(name, sizeof(name), "encoder.stages.%d.0.convs.%d", s, i);
      if (int r = add_conv(name, cur, i == 0 ? cur_s2d : ActView(), cur_c, f, a.kernels[s], stride, s, dst, s2d_out,
                           s < n - 1 ? a.strides[s + 1] : nullptr))
        return r;
      ConvStep& st = L.steps.back();
      if (fuse && last && s < n - 1) {  // skip producer: its scale / shift also fill the concat's combined table
        st.d_scale2 = comb_scale[s]; st.d_shift2 = comb_shift[s]; st.stride2 = 2 * f; st.off2 = f;
      }
      cur = feed_of_last(dst);
      cur_c = f;
      cur_s2d = ActView();
      // the copy exists when a pass (unfused) or a tensor-core epilogue (fused) writes it
      if (s2d_out && (fuse ? st.out_s2d != nullptr : true)) {
        const int* ns = a.strides[s + 1];
        const int phases = ns[0] * ns[1] * ns[2];
        cur_s2d.base = s2d_out; cur_s2d.groups_total = phases * f / 8; cur_s2d.group_off = 0;
        cur_s2d.groups = phases * f / 8;
        cur_s2d.D = dims[s][0] / ns[0]; cur_s2d.H = dims[s][1] / ns[1]; cur_s2d.W = dims[s][2] / ns[2];
      }
    }
  }
  // decoder
  for (int j = 0; j < n - 1; ++j) {
    const int s_below = n - 1 - j, s = n - 2 - j;
    const int cb = a.features[s_below], f = a.features[s];
    const int* st3 = a.strides[s_below];
    ConvStep up;
    up.is_tconv = true;
    char name[64];
    snprintf(name, sizeof(name), "decoder.transpconvs.%d", j);
    up.name = name;
    up.cin = cb; up.cout = f;
    for (int k = 0; k < 3; ++k) up.stride[k] = st3[k];
    up.src = cur.view;
    up.xf = cur.xf;
    up.out = view(cat[s], 2 * f / 8, 0, f / 8, s);
    const size_t nph = (size_t)st3[0] * st3[1] * st3[2];
    const HostTensor *w, *bi;
    if (int r = need(net, up.name + ".weight", &w, (size_t)cb * f * nph)) return r;
    if (int r = need(net, up.name + ".bias", &bi, f)) return r;
    const std::vector<float> wr = round_fp16(w->data);
    up.d_w = upload(net, up.name + ".weight", wr);
    up.d_bias = upload(net, up.name + ".bias", bi->data);
    if (!up.d_w || !up.d_bias) return BOA_ERR_CUDA;
    up.macs = (double)nph * cb * f * vox(s_below);
    macs += up.macs;
    up.kind = STEP_TCONV_SIMT;
    bool tc_ok = st3[2] == 2 && st3[1] == 2 && (st3[0] == 1 || st3[0] == 2) && cb % 16 == 0 && (nph * f) % 32 == 0;
    if (tc_ok) {
      ConvIO io;
      io.out = up.out;
      io.xf = cur.xf;
      TapsGeom geo;
      for (int k = 0; k < 3; ++k) { geo.ks[k] = st3[k]; geo.stride[k] = st3[k]; }
      up.taps = conv_taps_plan_create(TAPS_TCONV, geo, wr.data(), bi->data.data(), cb, f, cur.view, B, io, nullptr);
      if (!up.taps) return BOA_ERR_CUDA;
      up.kind = STEP_TCONV_TAPS;
    }
    L.steps.push_back(up);
    cur = Feed();
    cur.view = view(cat[s], 2 * f / 8, 0, 2 * f / 8, s);
    if (fuse) {
      cur.xf.scale = comb_scale[s]; cur.xf.shift = comb_shift[s]; cur.xf.channels = 2 * f;
      cur.xf.ident_groups = f / 8; cur.xf.slope = a.leaky_slope;
    }
    cur_c = 2 * f;
    for (int i = 0; i < a.n_conv_dec[j]; ++i) {
      const bool last = i == a.n_conv_dec[j] - 1;
      const int one[3] = {1, 1, 1};
      ActView dst;
      if (fuse) { dst = view(raw[s][raw_flip & 1], f / 8, 0, f / 8, s); ++raw_flip; }
      else dst = last ? view(outb[s], f / 8, 0, f / 8, s) : view(mid[s][i & 1], f / 8, 0, f / 8, s);
      snprintf(name, sizeof(name), "decoder.stages.%d.convs.%d", j, i);
      if (int r = add_conv(name, cur, ActView(), cur_c, f, a.kernels[s], one, s, dst, nullptr, nullptr)) return r;
      cur = feed_of_last(dst);
      cur_c = f;
    }
  }
  // head
  {
    char name[64];
    snprintf(name, sizeof(name), "decoder.seg_layers.%d", n - 2);
    const HostTensor *w, *bi;
    if (int r = need(net, std::string(name) + ".weight", &w, (size_t)a.num_classes * a.features[0])) return r;
    if (int r = need(net, std::string(name) + ".bias", &bi, a.num_classes)) return r;
    net->d_head_w = upload(net, std::string(name) + ".weight", round_fp16(w->data));
    net->d_head_b = upload(net, std::string(name) + ".bias", bi->data);
    if (!net->d_head_w || !net->d_head_b) return BOA_ERR_CUDA;
    L.head_src = cur.view;
    ConvStep& last = L.steps.back();
    BOA_REQUIRE(!last.is_tconv && last.cout == a.features[0], "boa_net_finalize: the decoder must end with a conv block");
    if (fuse || getenv("BOA_B200_NO_FUSE_HEAD") == nullptr) {
      // the head normalises the raw output of the last conv itself (free: it is HBM bound)
      L.head_src_raw = last.out;
      L.head_scale = last.d_scale; L.head_shift = last.d_shift;
      last.norm_pass = false;
    }
    macs += (double)a.num_classes * a.features[0] * vox(0);
  }
  net->macs_per_patch = (int64_t)macs;
  L.d_call = wsalloc<FwdCall>(net, li, 1);
  if (!L.d_call) return BOA_ERR_CUDA;
  return BOA_OK;
}

extern "C" int boa_net_finalize(boa_net* net) {
  BOA_REQUIRE(net && !net->finalized, "boa_net_finalize: bad state");
  BOA_CUDA(cudaSetDevice(net->device));
  for (int li = 0; li < net->n_lanes; ++li)
    if (int r = build_lane(net, li)) return r;
  Workspace* ws = net->ws;
  if (!ws->stream[0]) {
    for (int li = 0; li < MAX_LANES; ++li) {
      BOA_CUDA(cudaStreamCreateWithFlags(&ws->stream[li], cudaStreamNonBlocking));
      BOA_CUDA(cudaEventCreateWithFlags(&ws->join[li], cudaEventDisableTiming));
    }
    BOA_CUDA(cudaEventCreateWithFlags(&ws->head_done, cudaEventDisableTiming));
    BOA_CUDA(cudaEventCreateWithFlags(&ws->fork, cudaEventDisableTiming));
  }
  BOA_CUDA(cudaMallocHost(&net->h_call, sizeof(FwdCall) * boa_net::CALL_RING));
  for (int i = 0; i < boa_net::CALL_RING; ++i) BOA_CUDA(cudaEventCreateWithFlags(&net->call_ev[i], cudaEventDisableTiming));
  net->tensors.clear();
  net->finalized = true;
  return BOA_OK;
}

extern "C" int boa_net_forward_accumulate(boa_net* net, const float* d_vol, const int32_t* vol_shape,
                                          const int32_t* h_origins, int n_patches, const float* d_gaussian,
                                          float* d_logits_acc, void* stream) {
  BOA_REQUIRE(net && net->finalized, "boa_net_forward_accumulate: network not finalized");
  BOA_REQUIRE(d_vol && vol_shape && h_origins && d_gaussian && d_logits_acc, "boa_net_forward_accumulate: null");
  const boa_arch& a = net->arch;
  for (int p = 0; p < n_patches; ++p)
    for (int k = 0; k < 3; ++k)
      BOA_REQUIRE(h_origins[3 * p + k] >= 0 && h_origins[3 * p + k] + a.patch[k] <= vol_shape[k],
                  "boa_net_forward_accumulate: patch %d outside the volume", p);
  BOA_CUDA(cudaSetDevice(net->device));
  cudaStream_t user = static_cast<cudaStream_t>(stream);
  Workspace* ws = net->ws;
  const bool timing = net->timing;
  const bool lanes = net->n_lanes > 1 && !timing;
  if (lanes) {  // fork: both lane streams continue after everything enqueued on the caller's stream so far
    BOA_CUDA(cudaEventRecord(ws->fork, user));
    for (int li = 0; li < net->n_lanes; ++li) BOA_CUDA(cudaStreamWaitEvent(ws->stream[li], ws->fork, 0));
    ws->last_lane = -1;
  }
  int batch = 0;
  for (int p0 = 0; p0 < n_patches; p0 += net->B, ++batch) {
    const int nb = std::min(net->B, n_patches - p0);
    const int li = lanes ? (batch % net->n_lanes) : 0;
    Lane& L = net->lane[li];
    cudaStream_t s = lanes ? ws->stream[li] : user;
    // pinned staging slots are recycled round-robin: wait only for the copy that last used this slot
    const int slot = net->call_slot;
    net->call_slot = (slot + 1) % boa_net::CALL_RING;
    BOA_CUDA(cudaEventSynchronize(net->call_ev[slot]));
    FwdCall* c = net->h_call + slot;
    c->vol = d_vol; c->acc = d_logits_acc; c->gaussian = d_gaussian;
    c->d0 = vol_shape[0]; c->d1 = vol_shape[1]; c->d2 = vol_shape[2];
    c->n_valid = nb;
    for (int b = 0; b < MAX_BATCH; ++b)
      for (int k = 0; k < 3; ++k) c->origins[b][k] = b < nb ? h_origins[3 * (p0 + b) + k] : 0;
    BOA_CUDA(cudaMemcpyAsync(L.d_call, c, sizeof(FwdCall), cudaMemcpyHostToDevice, s));
    BOA_CUDA(cudaEventRecord(net->call_ev[slot], s));
    size_t f0 = 0, f1 = 0;
    if (timing) cudaEventRecord(next_event(net, &f0), s);
    const bool graph = net->use_graph && !timing;
    net->cur_nb = graph ? net->B : nb;  // plain launches skip the padding items of a partial batch entirely
    if (graph) {
      if (!L.g_body) {
        if (int r = capture_graph(&L.g_body, &L.launches_body, [&](cudaStream_t cs) { return run_front(net, L, cs); }))
          return r;
        if (int r = capture_graph(&L.g_head, &L.launches_head,
                                  [&](cudaStream_t cs) { return run_heads(net, L, net->B, nullptr, cs); }))
          return r;
      }
      BOA_CUDA(cudaGraphLaunch(L.g_body, s));
      count_launch(L.launches_body);
    } else if (int r = run_front(net, L, s)) {
      return r;
    }
    // the accumulation is ordered: this batch's heads run after the previous batch's heads (other lane)
    if (lanes && ws->last_lane >= 0 && ws->last_lane != li) BOA_CUDA(cudaStreamWaitEvent(s, ws->head_done, 0));
    if (graph) {
      BOA_CUDA(cudaGraphLaunch(L.g_head, s));
      count_launch(L.launches_head);
    } else if (int r = run_heads(net, L, nb, nullptr, s)) {
      return r;
    }
    if (lanes) {
      BOA_CUDA(cudaEventRecord(ws->head_done, s));
      ws->last_lane = li;
    }
    if (timing) {
      cudaEventRecord(next_event(net, &f1), s);
      net->fwd_spans.push_back({f0, f1});
    }
  }
  if (lanes) {  // join: the caller's stream continues after both lanes
    for (int li = 0; li < net->n_lanes; ++li) {
      BOA_CUDA(cudaEventRecord(ws->join[li], ws->stream[li]));
      BOA_CUDA(cudaStreamWaitEvent(user, ws->join[li], 0));
    }
  }
  return BOA_OK;
}

extern "C" int boa_net_forward_logits(boa_net* net, const float* d_patches, int n_patches, float* d_logits,
                                      void* stream) {
  BOA_REQUIRE(net && net->finalized, "boa_net_forward_logits: network not finalized");
  BOA_REQUIRE(d_patches && d_logits, "boa_net_forward_logits: null");
  const boa_arch& a = net->arch;
  BOA_CUDA(cudaSetDevice(net->device));
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  const size_t pv = (size_t)a.patch[0] * a.patch[1] * a.patch[2];
  for (int p0 = 0; p0 < n_patches; p0 += net->B) {
    const int nb = std::min(net->B, n_patches - p0);
    Lane& L = net->lane[0];
    net->cur_nb = nb;
    if (nb < net->B) BOA_CUDA(cudaMemsetAsync(L.d_patch, 0, (size_t)net->B * 16 * pv * sizeof(__half), s));
    if (net->input_mode == 2) {
      if (int r = launch_pack_patches_nb9(d_patches + (size_t)p0 * pv, nb, a.patch[0], a.patch[1], a.patch[2],
                                          L.d_patch, s))
        return r;
    } else if (net->input_mode == 1) {
      if (int r = launch_pack_patches_plain(d_patches + (size_t)p0 * pv, (size_t)nb * pv, L.d_patch, s)) return r;
    } else if (int r = launch_pack_patches(d_patches + (size_t)p0 * a.in_channels * pv, nb, a.in_channels,
                                           a.patch[0], a.patch[1], a.patch[2], L.d_patch, 2, s)) {
      return r;
    }
    if (int r = run_body(net, L, s)) return r;
    if (int r = run_heads(net, L, nb, d_logits + (size_t)p0 * a.num_classes * pv, s)) return r;
  }
  return BOA_OK;
}

extern "C" int64_t boa_net_macs_per_patch(const boa_net* net) { return net ? net->macs_per_patch : 0; }

extern "C" int boa_net_enable_timing(boa_net* net, int enable) {
  BOA_REQUIRE(net, "boa_net_enable_timing: null");
  net->timing = enable != 0;
  return BOA_OK;
}

extern "C" int boa_net_read_timing(boa_net* net, double* ms_convs, double* ms_total, int64_t* n_conv_launches,
                                   int reset) {
  BOA_REQUIRE(net, "boa_net_read_timing: null");
  BOA_CUDA(cudaSetDevice(net->device));
  BOA_CUDA(cudaDeviceSynchronize());
  double mc = 0, mt = 0;
  int64_t n_convs = 0;
  for (size_t i = 0; i < net->conv_spans.size(); ++i) {
    if (i < net->conv_span_info.size() && net->conv_span_info[i].first >= 6) continue;  // head launches
    const auto& sp = net->conv_spans[i];
    float ms = 0;
    cudaEventElapsedTime(&ms, net->ev[sp.first], net->ev[sp.second]);
    mc += ms;
    ++n_convs;
  }
  for (auto& sp : net->fwd_spans) {
    float ms = 0;
    cudaEventElapsedTime(&ms, net->ev[sp.first], net->ev[sp.second]);
    mt += ms;
  }
  if (ms_convs) *ms_convs = mc;
  if (ms_total) *ms_total = mt;
  if (n_conv_launches) *n_conv_launches = n_convs;
  if (reset) {
    net->conv_spans.clear();
    net->conv_span_info.clear();
    net->fwd_spans.clear();
    net->ev_used = 0;
  }
  return BOA_OK;
}

// Per kernel kind (StepKind order: fold, stride-2 taps, conv SIMT, transposed taps, transposed SIMT, first-layer SIMT):
// milliseconds, algorithmic FLOP and launches of the conv kernels recorded since the last reset while timing was
// enabled - i.e. measured INSIDE real forward_accumulate calls (single-lane schedule, so the event brackets are
// exclusive), under the clocks of a long step.
extern "C" int boa_net_read_timing_kinds(boa_net* net, int n_kinds, double* ms, double* flop, int64_t* launches,
                                         int reset) {
  BOA_REQUIRE(net && ms && flop && launches && n_kinds > 0, "boa_net_read_timing_kinds: bad argument");
  BOA_CUDA(cudaSetDevice(net->device));
  BOA_CUDA(cudaDeviceSynchronize());
  for (int k = 0; k < n_kinds; ++k) { ms[k] = 0; flop[k] = 0; launches[k] = 0; }
  for (size_t i = 0; i < net->conv_spans.size() && i < net->conv_span_info.size(); ++i) {
    const int k = net->conv_span_info[i].first;
    if (k < 0 || k >= n_kinds) continue;
    float t = 0;
    cudaEventElapsedTime(&t, net->ev[net->conv_spans[i].first], net->ev[net->conv_spans[i].second]);
    ms[k] += t;
    flop[k] += net->conv_span_info[i].second;
    ++launches[k];
  }
  if (reset) {
    net->conv_spans.clear();
    net->conv_span_info.clear();
    net->fwd_spans.clear();
    net->ev_used = 0;
  }
  return BOA_OK;
}

// Per-layer description for profiling / DESIGN.md: writes up to `cap` entries, returns the number of steps.
extern "C" int boa_net_describe(const boa_net* net, int cap, int32_t* kinds, double* macs, char* names, int name_stride) {
  if (!net) return 0;
  const std::vector<ConvStep>& steps = net->lane[0].steps;
  const int n = (int)steps.size();
  for (int i = 0; i < n && i < cap; ++i) {
    if (kinds) kinds[i] = (int)steps[i].kind;
    if (macs) macs[i] = steps[i].macs;
    if (names) snprintf(names + (size_t)i * name_stride, name_stride, "%s", steps[i].name.c_str());
  }
  return n;
}

// Time each step of one forward separately (bench / profiling): ms[i] = device time of step i's conv kernel.
extern "C" int boa_net_time_layers(boa_net* net, int cap, float* ms, void* stream) {
  BOA_REQUIRE(net && net->finalized && ms, "boa_net_time_layers: bad argument");
  BOA_CUDA(cudaSetDevice(net->device));
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  const bool was = net->timing;
  BOA_CUDA(cudaDeviceSynchronize());
  net->conv_spans.clear();
  net->conv_span_info.clear();
  net->fwd_spans.clear();
  net->ev_used = 0;
  net->timing = true;
  net->cur_nb = net->B;
  int r = run_body(net, net->lane[0], s);
  net->timing = was;
  if (r) return r;
  BOA_CUDA(cudaDeviceSynchronize());
  for (size_t i = 0; i < net->conv_spans.size() && (int)i < cap; ++i)
    cudaEventElapsedTime(&ms[i], net->ev[net->conv_spans[i].first], net->ev[net->conv_spans[i].second]);
  net->conv_spans.clear();
  net->conv_span_info.clear();
  net->ev_used = 0;
  return BOA_OK;
}

extern "C" void boa_net_destroy(boa_net* net) {
  if (!net) return;
  cudaSetDevice(net->device);
  cudaDeviceSynchronize();
  destroy_graphs(net);
  for (Lane& L : net->lane)
    for (ConvStep& st : L.steps) {
      if (st.fold) conv_mma_plan_destroy(st.fold);
      if (st.taps) conv_taps_plan_destroy(st.taps);
    }
  for (void* p : net->allocs) cudaFree(p);
  if (--net->ws->refs == 0) {
    for (auto& bufs : net->ws->bufs)
      for (auto& b : bufs) cudaFree(b.first);
    for (int li = 0; li < MAX_LANES; ++li) {
      if (net->ws->stream[li]) cudaStreamDestroy(net->ws->stream[li]);
      if (net->ws->join[li]) cudaEventDestroy(net->ws->join[li]);
    }
    if (net->ws->head_done) cudaEventDestroy(net->ws->head_done);
    if (net->ws->fork) cudaEventDestroy(net->ws->fork);
    delete net->ws;
  }
  for (cudaEvent_t e : net->ev) cudaEventDestroy(e);
  if (net->h_call) cudaFreeHost(net->h_call);
  for (cudaEvent_t e : net->call_ev)
    if (e) cudaEventDestroy(e);
  delete net;
}
