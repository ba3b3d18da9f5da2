// C ABI of libgnina_b200.so (see include/gnina_b200.h for the reference interfaces each entry replaces).
#include <algorithm>
#include <cmath>
#include <cstring>
#include <fstream>
#include <mutex>
#include "gb_internal.h"
#include "gb_tc.h"

namespace gb {
static thread_local std::string g_last_error;
void set_last_error(const std::string& m) { g_last_error = m; }
void prefer_many_hw_queues() {
  static const int once = (setenv("CUDA_DEVICE_MAX_CONNECTIONS", "32", 0), 0);
  (void)once;
}

void Profiler::begin(const char* name, cudaStream_t s) {
  int id = -1;
  for (size_t i = 0; i < names.size(); i++)
    if (names[i] == name) id = (int)i;
  if (id < 0) { names.push_back(name); total_ms.push_back(0); count.push_back(0); id = (int)names.size() - 1; }
  auto get = [&]() { cudaEvent_t e; if (pool.empty()) { cudaEventCreate(&e); } else { e = pool.back(); pool.pop_back(); } return e; };
  cur = id; cur_a = get();
  cudaEventRecord(cur_a, s);
}
void Profiler::end(cudaStream_t s) {
  if (cur < 0) return;
  cudaEvent_t b; if (pool.empty()) cudaEventCreate(&b); else { b = pool.back(); pool.pop_back(); }
  cudaEventRecord(b, s);
  pending.push_back({cur, cur_a, b});
  cur = -1;
}
void Profiler::resolve() {
  for (auto& p : pending) {
    float ms = 0;
    if (cudaEventElapsedTime(&ms, p.a, p.b) == cudaSuccess) { total_ms[p.id] += ms; count[p.id]++; }
    pool.push_back(p.a); pool.push_back(p.b);
  }
  pending.clear();
}
void Profiler::reset() { resolve(); for (auto& t : total_ms) t = 0; for (auto& c : count) c = 0; }
Profiler::~Profiler() { for (auto& p : pending) { cudaEventDestroy(p.a); cudaEventDestroy(p.b); } for (auto e : pool) cudaEventDestroy(e); }

template <typename T>
struct DevBuf {
  T* p = nullptr;
  size_t cap = 0;
  void ensure(size_t n) {
    if (cap >= n) return;
    n = std::max<size_t>(std::max<size_t>(n, 4096), cap + cap / 2);  // floor + geometric growth: few device-wide syncs
    if (p) cudaFree(p);
    p = nullptr;
    GB_CUDA(cudaMalloc(&p, n * sizeof(T)));
    cap = n;
  }
  ~DevBuf() { if (p) cudaFree(p); }
};
template <typename T>
struct PinBuf {
  T* p = nullptr;
  size_t cap = 0;
  void ensure(size_t n) {
    if (cap >= n) return;
    n = std::max<size_t>(std::max<size_t>(n, 4096), cap + cap / 2);
    if (p) cudaFreeHost(p);
    p = nullptr;
    GB_CUDA(cudaHostAlloc(&p, n * sizeof(T), cudaHostAllocDefault));
    cap = n;
  }
  ~PinBuf() { if (p) cudaFreeHost(p); }
};

// Models that voxelise identically (same type maps and grid metadata) share typed atoms, pose lists and grids.
struct GridGroup {
  GridSig sig;
  TypeMap rec, lig;
  int n_channels = 0;
  std::vector<int> model_idx;
  // receptor on device (typed atoms only, stably sorted by channel)
  DevBuf<float4> rec_xyzr;
  DevBuf<int> rec_ch;
  int n_rec = 0;
  std::vector<int> h_rec_src;   // typed receptor atom -> index in the caller's receptor arrays
  DevBuf<int> rec_off;          // {0, n_rec}: the receptor as one "pose" for the atom-gradient kernels
  DevBuf<float> rec_grad;       // [n_rec][3], accumulated over the models of the group
  PinBuf<float> h_rec_grad;
  // staged ligand atoms
  PinBuf<float4> h_lig_xyzr;
  PinBuf<int> h_lig_ch, h_lig_off, h_lig_src;  // h_lig_src: staged atom -> index in the caller's arrays
  DevBuf<float> lig_grad;   // [staged atoms][3], accumulated over the models of the group
  PinBuf<float> h_lig_grad;
  int n_staged_atoms = 0;
  DevBuf<float4> lig_xyzr;
  DevBuf<int> lig_ch, lig_off;
  int max_pose_atoms = 0;
  // per-chunk workspaces
  DevBuf<float4> list_xyzr;
  DevBuf<int> list_ch, list_n;
  DevBuf<float> grid;       // fp32 reference layout [chunk][C][N^3]
  TcGridWorkspace tc_grid;  // fused pooled fp16 layout (fast path)
};
}  // namespace gb

using namespace gb;

struct gb_model {
  Model* m;
};

struct gb_cnn {
  int device = 0;
  cudaStream_t stream = nullptr;
  cudaStream_t aux = nullptr;   // voxeliser of the next chunk overlaps the network of the current one
  cudaEvent_t ev_fork = nullptr;
  cudaEvent_t ev_staged = nullptr;   // deferred staging: the H2D copies of a chunk (on aux) are done
  std::vector<Model*> models;
  std::vector<std::unique_ptr<GridGroup>> groups;
  std::vector<int> model_group;
  int precision = GB_PRECISION_FP32;
  int max_batch = 0;  // 0 = per-precision default
  // deferred staging (gb_cnn_score_batch): the caller's arrays, valid for the duration of that call only
  struct LazyInput { const float* xyz = nullptr; const int32_t* type = nullptr; const int32_t* off = nullptr; const float* centers = nullptr; int n = 0, upto = 0; };
  LazyInput lazy;
  int overlap = 0;    // 1: voxelise chunk i+1 on the aux stream while the network of chunk i runs
  int cnn_rotation = 0;            // --cnn_rotation: evaluations per model; 0/1 = the unrotated pose only
  uint32_t rotation_seed = 0;      // --seed
  PinBuf<float> h_rot;             // [R][n_staged][9]
  DevBuf<float> d_rot;
  std::vector<float> rec_xyz;
  std::vector<int32_t> rec_type;
  // staged poses
  int n_staged = 0;
  PinBuf<float> h_centers;
  DevBuf<float> d_centers;
  DevBuf<float> d_out3;                  // [chunk][3]
  DevBuf<float> d_pose, d_aff, d_loss;   // [M][n_staged]
  DevBuf<float> d_final;                 // [4][n_staged]
  PinBuf<float> h_final;
  Fp32Workspace ws32;
  Fp32GradWorkspace ws_grad;
  DevBuf<float> d_dgrid;
  DevBuf<float> tmp_grad, tmp_rec;   // per-chunk atom gradients of gb_cnn_score_grad
  int n_input_atoms = 0;
  TcWorkspace ws_tc;
  int64_t launches = 0;
  Profiler prof;
  ~gb_cnn() {
    if (stream) cudaStreamDestroy(stream);
    if (aux) cudaStreamDestroy(aux);
    if (ev_fork) cudaEventDestroy(ev_fork);
    if (ev_staged) cudaEventDestroy(ev_staged);
    for (Model* m : models)
      if (--m->refs == 0) delete m;
  }
};

#define GB_API_BEGIN try {
#define GB_API_END                          \
  }                                         \
  catch (const gb::Error& e) {              \
    gb::set_last_error(e.what());           \
    return e.code;                          \
  }                                         \
  catch (const std::exception& e) {         \
    gb::set_last_error(e.what());           \
    return GB_ERR_INTERNAL;                 \
  }                                         \
  return GB_OK;

static void require_device() {
  prefer_many_hw_queues();
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess || n == 0) {
    cudaGetLastError();
    throw Error(GB_ERR_NO_DEVICE, "gnina_b200: no CUDA device visible — this library has no CPU fallback");
  }
}

extern "C" {

const char* gb_last_error(void) { return g_last_error.c_str(); }
const char* gb_version(void) { return "gnina_b200 0.1 (sm_100a)"; }

int gb_device_count(void) {
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
  return n;
}

int gb_initialize_cuda(int device) {
  prefer_many_hw_queues();
  cudaError_t e = cudaSetDevice(device);
  if (e != cudaSuccess) { cudaGetLastError(); return (int)e; }
  cudaDeviceProp prop;
  e = cudaGetDeviceProperties(&prop, device);
  if (e != cudaSuccess) return (int)e;
  if (prop.computeMode == cudaComputeModeProhibited) return -1;
  return 0;
}

int gb_model_load_mem(const void* data, size_t nbytes, int device, gb_model** out) {
  GB_API_BEGIN
  GB_CHECK(out && data, "null argument");
  require_device();
  *out = new gb_model{load_model_from_memory(data, nbytes, device, "<memory>")};
  GB_API_END
}

int gb_model_load(const char* path, int device, gb_model** out) {
  GB_API_BEGIN
  GB_CHECK(out && path, "null argument");
  std::ifstream f(path, std::ios::binary);
  if (!f) throw Error(GB_ERR_USAGE, std::string("Could not open file ") + path);  // cnn_torch_scorer.cpp:86-87
  std::vector<char> raw((std::istreambuf_iterator<char>(f)), std::istreambuf_iterator<char>());
  require_device();
  *out = new gb_model{load_model_from_memory(raw.data(), raw.size(), device, path)};
  GB_API_END
}

int gb_model_get_info(const gb_model* m, gb_model_info* info) {
  GB_API_BEGIN
  GB_CHECK(m && info, "null argument");
  const Model& M = *m->m;
  memset(info, 0, sizeof(*info));
  info->arch = M.arch;
  info->n_rec_channels = M.rec.n_channels;
  info->n_lig_channels = M.lig.n_channels;
  info->grid_points = M.npts;
  info->resolution = M.resolution;
  info->dimension = M.dimension;
  info->radius_scaling = M.radius_scaling;
  info->apply_logistic_loss = M.apply_logistic_loss;
  info->skip_softmax = M.skip_softmax;
  strncpy(info->name, M.name.c_str(), sizeof(info->name) - 1);
  GB_API_END
}

void gb_model_release(gb_model* m) {
  if (!m) return;
  if (--m->m->refs == 0) delete m->m;
  delete m;
}

int gb_model_type_atoms(const gb_model* m, int is_ligand, const int32_t* smina_type, int n, int32_t* channel,
                        float* radius) {
  GB_API_BEGIN
  GB_CHECK(m && smina_type && channel && radius, "null argument");
  const TypeMap& tm = is_ligand ? m->m->lig : m->m->rec;
  const int off = is_ligand ? m->m->rec.n_channels : 0;
  for (int i = 0; i < n; i++) {
    const int t = smina_type[i];
    const bool ok = t >= 0 && t < kNumSminaTypes;
    const int c = ok ? tm.t2c[t] : -1;
    channel[i] = c < 0 ? -1 : c + off;
    radius[i] = ok ? kSminaXsRadius[t] : 0.f;
  }
  GB_API_END
}

static void build_groups(gb_cnn* h) {
  h->groups.clear();
  h->model_group.assign(h->models.size(), -1);
  for (size_t i = 0; i < h->models.size(); i++) {
    const Model& M = *h->models[i];
    GridSig sig{M.recmap, M.ligmap, M.resolution, M.dimension, M.radius_scaling};
    int g = -1;
    for (size_t k = 0; k < h->groups.size(); k++)
      if (h->groups[k]->sig == sig) g = (int)k;
    if (g < 0) {
      auto G = std::make_unique<GridGroup>();
      G->sig = sig; G->rec = M.rec; G->lig = M.lig; G->n_channels = M.n_channels;
      h->groups.push_back(std::move(G));
      g = (int)h->groups.size() - 1;
    }
    h->groups[g]->model_idx.push_back((int)i);
    h->model_group[i] = g;
  }
}

int gb_cnn_create(gb_model* const* models, int n_models, int device, gb_cnn** out) {
  GB_API_BEGIN
  GB_CHECK(out && models && n_models > 0, "gb_cnn_create needs at least one model");
  require_device();
  GB_CUDA(cudaSetDevice(device));
  std::unique_ptr<gb_cnn> h(new gb_cnn);
  h->device = device;
  for (int i = 0; i < n_models; i++) {
    GB_CHECK(models[i] && models[i]->m->device == device, "model loaded on a different device");
    models[i]->m->refs++;
    h->models.push_back(models[i]->m);
  }
  GB_CUDA(cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking));
  GB_CUDA(cudaStreamCreateWithFlags(&h->aux, cudaStreamNonBlocking));
  GB_CUDA(cudaEventCreateWithFlags(&h->ev_fork, cudaEventDisableTiming));
  GB_CUDA(cudaEventCreateWithFlags(&h->ev_staged, cudaEventDisableTiming));
  build_groups(h.get());
  h->precision = GB_PRECISION_FP16_TC;
  for (Model* m : h->models)
    if (!tc_supported(*m)) h->precision = GB_PRECISION_FP32;
  *out = h.release();
  GB_API_END
}

int gb_cnn_set_receptor(gb_cnn* h, const float* xyz, const int32_t* smina_type, int n);

int gb_cnn_clone(const gb_cnn* src, gb_cnn** out) {
  GB_API_BEGIN
  GB_CHECK(src && out, "null argument");
  GB_CUDA(cudaSetDevice(src->device));
  std::unique_ptr<gb_cnn> h(new gb_cnn);
  h->device = src->device;
  for (Model* m : src->models) { m->refs++; h->models.push_back(m); }
  GB_CUDA(cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking));
  GB_CUDA(cudaStreamCreateWithFlags(&h->aux, cudaStreamNonBlocking));
  GB_CUDA(cudaEventCreateWithFlags(&h->ev_fork, cudaEventDisableTiming));
  GB_CUDA(cudaEventCreateWithFlags(&h->ev_staged, cudaEventDisableTiming));
  build_groups(h.get());
  h->precision = src->precision;
  h->max_batch = src->max_batch;
  h->cnn_rotation = src->cnn_rotation;
  h->rotation_seed = src->rotation_seed;
  h->overlap = src->overlap;
  h->prof.on = src->prof.on;
  *out = nullptr;
  if (!src->rec_type.empty()) {
    const int rc = gb_cnn_set_receptor(h.get(), src->rec_xyz.data(), src->rec_type.data(), (int)src->rec_type.size());
    if (rc != GB_OK) return rc;   // the half-built clone is destroyed with h; *out stays null
  }
  *out = h.release();
  GB_API_END
}

void gb_cnn_destroy(gb_cnn* h) {
  if (!h) return;
  cudaSetDevice(h->device);
  if (h->stream) cudaStreamSynchronize(h->stream);
  delete h;
}

int gb_cnn_num_models(const gb_cnn* h) { return h ? (int)h->models.size() : 0; }

int gb_cnn_set_option(gb_cnn* h, const char* key, double value) {
  GB_API_BEGIN
  GB_CHECK(h && key, "null argument");
  const std::string k(key);
  if (k == "precision") {
    const int p = (int)value;
    if (p != GB_PRECISION_FP32 && p != GB_PRECISION_FP16_TC) throw Error(GB_ERR_USAGE, "unknown precision");
    if (p == GB_PRECISION_FP16_TC)
      for (Model* m : h->models)
        if (!tc_supported(*m)) throw Error(GB_ERR_USAGE, "model " + m->name + " has no tensor-core path yet");
    h->precision = p;
  } else if (k == "overlap") {
    h->overlap = value != 0;
  } else if (k == "profile") {
    h->prof.on = value != 0;
  } else if (k == "cnn_rotation") {
    GB_CHECK(value >= 0 && value <= 24, "cnn_rotation out of range (0..24, as gnina's option)");
    h->cnn_rotation = (int)value;
  } else if (k == "rotation_seed") {
    h->rotation_seed = (uint32_t)value;
  } else if (k == "max_batch") {
    GB_CHECK(value >= 0 && value <= 65536, "max_batch out of range");
    h->max_batch = (int)value;
  } else {
    throw Error(GB_ERR_USAGE, "unknown option " + k);
  }
  GB_API_END
}

double gb_cnn_get_option(const gb_cnn* h, const char* key) {
  if (!h || !key) return NAN;
  const std::string k(key);
  if (k == "precision") return h->precision;
  if (k == "max_batch") return h->max_batch;
  if (k == "overlap") return h->overlap;
  if (k == "cnn_rotation") return h->cnn_rotation;
  if (k == "rotation_seed") return h->rotation_seed;
  return NAN;
}

int gb_cnn_set_receptor(gb_cnn* h, const float* xyz, const int32_t* smina_type, int n) {
  GB_API_BEGIN
  GB_CHECK(h && (n == 0 || (xyz && smina_type)) && n >= 0, "bad receptor arguments");
  GB_CUDA(cudaSetDevice(h->device));
  h->rec_xyz.assign(xyz, xyz + 3 * (size_t)n);
  h->rec_type.assign(smina_type, smina_type + n);
  for (auto& Gp : h->groups) {
    GridGroup& G = *Gp;
    // make_coordset (torch_model.cpp:120-142) once; untyped atoms (channel -1: hydrogens) never contribute
    std::vector<std::vector<int>> by_ch(G.rec.n_channels);
    for (int i = 0; i < n; i++) {
      const int t = smina_type[i];
      if (t < 0 || t >= kNumSminaTypes) continue;
      const int c = G.rec.t2c[t];
      if (c >= 0) by_ch[c].push_back(i);
    }
    std::vector<float4> a;
    std::vector<int> ch;
    G.h_rec_src.clear();
    for (int c = 0; c < G.rec.n_channels; c++)
      for (int i : by_ch[c]) {
        a.push_back(make_float4(xyz[3 * i], xyz[3 * i + 1], xyz[3 * i + 2],
                                kSminaXsRadius[smina_type[i]] * G.sig.radius_scaling));
        ch.push_back(c);
        G.h_rec_src.push_back(i);
      }
    G.n_rec = (int)a.size();
    {
      const int off[2] = {0, G.n_rec};
      G.rec_off.ensure(2);
      GB_CUDA(cudaMemcpyAsync(G.rec_off.p, off, sizeof(off), cudaMemcpyHostToDevice, h->stream));
      GB_CUDA(cudaStreamSynchronize(h->stream));
    }
    G.rec_xyzr.ensure(a.size());
    G.rec_ch.ensure(ch.size());
    if (G.n_rec) {
      GB_CUDA(cudaMemcpyAsync(G.rec_xyzr.p, a.data(), a.size() * sizeof(float4), cudaMemcpyHostToDevice, h->stream));
      GB_CUDA(cudaMemcpyAsync(G.rec_ch.p, ch.data(), ch.size() * sizeof(int), cudaMemcpyHostToDevice, h->stream));
      GB_CUDA(cudaStreamSynchronize(h->stream));
    }
  }
  GB_API_END
}

// Staging = typing + stable counting sort by channel of the ligand atoms on the host, pose centres, H2D copies.  It is done in
// RANGES of poses: gb_cnn_stage_poses stages everything at once; gb_cnn_score_batch defers it (h->lazy) so that gb_cnn_run_staged
// stages the poses of chunk i + 1 while the device works on chunk i -- for 10 k poses the host loop is 1.4 ms, 9 % of the
// end-to-end step when it runs in front of the first kernel.
static void stage_prepare(gb_cnn* h, const float* lig_xyz, const int32_t* lig_type, const int32_t* pose_offsets, int n_poses,
                          const float* centers) {
  GB_CHECK(h && n_poses >= 0 && (n_poses == 0 || (lig_xyz && lig_type && pose_offsets)), "bad pose arguments");
  GB_CUDA(cudaSetDevice(h->device));
  // the previous batch may still be reading the pinned staging buffers
  GB_CUDA(cudaStreamSynchronize(h->stream));
  h->n_staged = 0;  // nothing is staged until every check below has passed
  h->n_input_atoms = 0;
  h->lazy = gb_cnn::LazyInput();
  if (n_poses == 0) return;
  const int total = pose_offsets[n_poses];
  GB_CHECK(pose_offsets[0] == 0 && total >= 0, "pose_offsets must start at 0");
  int max_raw = 0;
  for (int p = 0; p < n_poses; p++) {
    GB_CHECK(pose_offsets[p + 1] >= pose_offsets[p], "pose_offsets must be non-decreasing");
    max_raw = std::max(max_raw, pose_offsets[p + 1] - pose_offsets[p]);
  }
  h->h_centers.ensure(3 * (size_t)n_poses);
  h->d_centers.ensure(3 * (size_t)n_poses);
  for (auto& Gp : h->groups) {
    GridGroup& G = *Gp;
    G.h_lig_xyzr.ensure((size_t)total);
    G.h_lig_ch.ensure((size_t)total);
    G.h_lig_off.ensure((size_t)n_poses + 1);
    G.h_lig_src.ensure((size_t)total);
    G.lig_xyzr.ensure((size_t)total);
    G.lig_ch.ensure((size_t)total);
    G.lig_off.ensure((size_t)n_poses + 1);
    G.n_staged_atoms = 0;            // running write offset while ranges are staged, the total afterwards
    G.max_pose_atoms = max_raw;      // capacity bound for the per-pose atom lists (typed atoms <= atoms passed)
  }
  h->lazy.xyz = lig_xyz; h->lazy.type = lig_type; h->lazy.off = pose_offsets; h->lazy.centers = centers;
  h->lazy.n = n_poses; h->lazy.upto = 0;
  h->n_staged = n_poses;
  h->n_input_atoms = total;
}

// stage poses [h->lazy.upto, p1): host work, then their H2D copies on stream cs
static void stage_range(gb_cnn* h, int p1, cudaStream_t cs) {
  gb_cnn::LazyInput& L = h->lazy;
  const int p0 = L.upto;
  if (p1 > L.n) p1 = L.n;
  if (p1 <= p0) return;
  const float* lig_xyz = L.xyz;
  const int32_t* lig_type = L.type;
  const int32_t* pose_offsets = L.off;
  for (int p = p0; p < p1; p++) {
    const int b = pose_offsets[p], e = pose_offsets[p + 1];
    if (L.centers) {
      for (int d = 0; d < 3; d++) h->h_centers.p[3 * p + d] = L.centers[3 * p + d];
    } else {
      // CoordinateSet::center(): float mean over ALL ligand atoms passed (torch_model.cpp:163-166)
      float sx = 0, sy = 0, sz = 0;
      for (int i = b; i < e; i++) { sx += lig_xyz[3 * i]; sy += lig_xyz[3 * i + 1]; sz += lig_xyz[3 * i + 2]; }
      const int cnt = e - b;
      if (cnt > 0) { sx /= cnt; sy /= cnt; sz /= cnt; }
      h->h_centers.p[3 * p] = sx; h->h_centers.p[3 * p + 1] = sy; h->h_centers.p[3 * p + 2] = sz;
    }
  }
  GB_CUDA(cudaMemcpyAsync(h->d_centers.p + 3 * (size_t)p0, h->h_centers.p + 3 * (size_t)p0, 3 * (size_t)(p1 - p0) * sizeof(float),
                          cudaMemcpyHostToDevice, cs));
  for (auto& Gp : h->groups) {
    GridGroup& G = *Gp;
    const int nrc = G.rec.n_channels, nlc = G.lig.n_channels;
    const int w0 = G.n_staged_atoms;
    int w = w0;
    int cnt[64], start[64];
    for (int p = p0; p < p1; p++) {
      const int b = pose_offsets[p], e = pose_offsets[p + 1];
      G.h_lig_off.p[p] = w;
      // counting sort by channel (stable): keeps the per-channel summation order of the reference
      for (int c = 0; c < nlc; c++) cnt[c] = 0;
      for (int i = b; i < e; i++) {
        const int t = lig_type[i];
        const int c = (t >= 0 && t < kNumSminaTypes) ? G.lig.t2c[t] : -1;
        if (c >= 0) cnt[c]++;
      }
      int s = w;
      for (int c = 0; c < nlc; c++) { start[c] = s; s += cnt[c]; }
      for (int i = b; i < e; i++) {
        const int t = lig_type[i];
        const int c = (t >= 0 && t < kNumSminaTypes) ? G.lig.t2c[t] : -1;
        if (c < 0) continue;
        const int dst = start[c]++;
        G.h_lig_xyzr.p[dst] = make_float4(lig_xyz[3 * i], lig_xyz[3 * i + 1], lig_xyz[3 * i + 2],
                                          kSminaXsRadius[t] * G.sig.radius_scaling);
        G.h_lig_ch.p[dst] = nrc + c;
        G.h_lig_src.p[dst] = i;
      }
      w = s;
    }
    G.h_lig_off.p[p1] = w;   // end of the range = offset of the next pose (rewritten, identically, by the next range)
    G.n_staged_atoms = w;
    if (w > w0) {
      GB_CUDA(cudaMemcpyAsync(G.lig_xyzr.p + w0, G.h_lig_xyzr.p + w0, (size_t)(w - w0) * sizeof(float4), cudaMemcpyHostToDevice, cs));
      GB_CUDA(cudaMemcpyAsync(G.lig_ch.p + w0, G.h_lig_ch.p + w0, (size_t)(w - w0) * sizeof(int), cudaMemcpyHostToDevice, cs));
    }
    // offsets p0 .. p1; entry p0 of a later range was already copied as the previous range's end (the chunk before may be reading it)
    const int o0 = p0 > 0 ? p0 + 1 : 0;
    GB_CUDA(cudaMemcpyAsync(G.lig_off.p + o0, G.h_lig_off.p + o0, ((size_t)(p1 - o0) + 1) * sizeof(int), cudaMemcpyHostToDevice, cs));
  }
  L.upto = p1;
}

int gb_cnn_stage_poses(gb_cnn* h, const float* lig_xyz, const int32_t* lig_type, const int32_t* pose_offsets,
                       int n_poses, const float* centers) {
  GB_API_BEGIN
  stage_prepare(h, lig_xyz, lig_type, pose_offsets, n_poses, centers);
  stage_range(h, n_poses, h->stream);
  h->lazy = gb_cnn::LazyInput();   // the caller's arrays are not referenced after this call
  GB_API_END
}

static int chunk_size(const gb_cnn* h) {
  if (h->max_batch > 0) return h->max_batch;
  // fast path: 4096 poses per chunk (workspaces grow to what a call actually needs, ~3.5 MB per pose); r4a: per-launch ramps
  // cost 1.6 % at 5000 and 3.3 % at 10000 poses per chunk relative to 2048
  return h->precision == GB_PRECISION_FP32 ? 16 : 4096;
}

// voxelise poses [p0, p0+nb) of group G into the fp32 reference layout
static void voxelize_chunk_f32(gb_cnn* h, GridGroup& G, int p0, int nb, const float* rot = nullptr) {
  const int cap = G.n_rec + G.max_pose_atoms;
  G.list_xyzr.ensure((size_t)nb * std::max(cap, 1));
  G.list_ch.ensure((size_t)nb * std::max(cap, 1));
  G.list_n.ensure(nb);
  const int npts = (int)std::lround(G.sig.dimension / G.sig.resolution) + 1;
  G.grid.ensure((size_t)nb * G.n_channels * npts * npts * npts);
  {
  ProfScope ps(&h->prof, "f32_build_pose_lists", h->stream);
  launch_build_pose_lists(G.rec_xyzr.p, G.rec_ch.p, G.n_rec, G.lig_xyzr.p, G.lig_ch.p, G.lig_off.p + p0,
                          h->d_centers.p + 3 * (size_t)p0, nb, G.sig.dimension / 2.f, std::max(cap, 1), G.list_xyzr.p,
                          G.list_ch.p, G.list_n.p, h->stream, rot);
  }
  ProfScope ps2(&h->prof, "f32_voxelize", h->stream);
  launch_voxelize_f32(G.list_xyzr.p, G.list_ch.p, G.list_n.p, std::max(cap, 1), h->d_centers.p + 3 * (size_t)p0, nb,
                      G.n_channels, npts, G.sig.resolution, G.sig.dimension, G.grid.p, h->stream);
  h->launches += 2;
}

// G3: rotation r >= 1 of staged pose p -- a uniformly random rotation (Shoemake's quaternion construction) from a
// counter-based generator keyed by (seed, r, p).  The reference draws from libmolgrid's generator (seeded per model,
// cnn_torch_scorer.cpp:131-132), which cannot be reproduced here: the MECHANISM is pinned (tests rotate the inputs
// on the host with the matrix returned by gb_cnn_get_rotation), the random stream is not.
static uint32_t rot_hash(uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
  uint32_t x = a * 0x9E3779B1u ^ (b + 0x7F4A7C15u) * 0x85EBCA77u ^ (c + 0x165667B1u) * 0xC2B2AE3Du ^ (d + 1u) * 0x27D4EB2Fu;
  x ^= x >> 15; x *= 0x2C1B3C6Du; x ^= x >> 12; x *= 0x297A2D39u; x ^= x >> 15;
  return x;
}
static void rotation_matrix(uint32_t seed, int r, int p, float* R) {
  if (r <= 0) { const float I[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1}; for (int k = 0; k < 9; k++) R[k] = I[k]; return; }
  auto u = [&](uint32_t k) { return (rot_hash(seed, (uint32_t)r, (uint32_t)p, k) >> 8) * (1.0 / 16777216.0); };
  const double u1 = u(0), u2 = u(1), u3 = u(2), pi2 = 6.283185307179586;
  const double a = std::sqrt(1 - u1), b = std::sqrt(u1);
  const double qx = a * std::sin(pi2 * u2), qy = a * std::cos(pi2 * u2), qz = b * std::sin(pi2 * u3), qw = b * std::cos(pi2 * u3);
  const double m[9] = {1 - 2 * (qy * qy + qz * qz), 2 * (qx * qy - qz * qw),     2 * (qx * qz + qy * qw),
                       2 * (qx * qy + qz * qw),     1 - 2 * (qx * qx + qz * qz), 2 * (qy * qz - qx * qw),
                       2 * (qx * qz - qy * qw),     2 * (qy * qz + qx * qw),     1 - 2 * (qx * qx + qy * qy)};
  for (int k = 0; k < 9; k++) R[k] = (float)m[k];
}
static int n_rotations(const gb_cnn* h) { return std::max(1, h->cnn_rotation); }

// matrices of every (rotation, staged pose) -> h->d_rot [R][n][9]
static void upload_rotations(gb_cnn* h) {
  const int R = n_rotations(h), n = h->n_staged;
  if (R <= 1 || n == 0) return;
  h->h_rot.ensure((size_t)R * n * 9);
  h->d_rot.ensure((size_t)R * n * 9);
  for (int r = 0; r < R; r++)
    for (int p = 0; p < n; p++) rotation_matrix(h->rotation_seed, r, p, h->h_rot.p + ((size_t)r * n + p) * 9);
  GB_CUDA(cudaMemcpyAsync(h->d_rot.p, h->h_rot.p, (size_t)R * n * 9 * sizeof(float), cudaMemcpyHostToDevice, h->stream));
}

int gb_cnn_get_rotation(const gb_cnn* h, int rotation, int pose, float* matrix9) {
  GB_API_BEGIN
  GB_CHECK(h && matrix9 && rotation >= 0 && rotation < n_rotations(h) && pose >= 0, "bad rotation query");
  rotation_matrix(h->rotation_seed, rotation, pose, matrix9);
  GB_API_END
}

int gb_cnn_run_staged(gb_cnn* h) {
  GB_API_BEGIN
  GB_CHECK(h, "null handle");
  GB_CUDA(cudaSetDevice(h->device));
  const int n = h->n_staged, M0 = (int)h->models.size();
  if (n == 0) return GB_OK;
  // every model is evaluated on R rotations of every pose (cnn_torch_scorer.cpp:127-163); the (model, rotation)
  // evaluations are the ensemble's members: slot mi * R + r
  const int R = n_rotations(h), M = M0 * R;
  upload_rotations(h);
  h->d_pose.ensure((size_t)M * n);
  h->d_aff.ensure((size_t)M * n);
  h->d_loss.ensure((size_t)M * n);
  h->d_final.ensure(4 * (size_t)n);
  const int chunk = chunk_size(h);
  h->d_out3.ensure(3 * (size_t)chunk);
  // the auxiliary stream starts after everything already queued on the main stream (staging copies, timing events)
  GB_CUDA(cudaEventRecord(h->ev_fork, h->stream));
  GB_CUDA(cudaStreamWaitEvent(h->aux, h->ev_fork, 0));
  // deferred staging: the poses of chunk i + 1 are staged -- copies on the auxiliary stream -- while the device works on chunk i;
  // only the first chunk's staging is in front of the first kernel.  (r4c, e2e for 10 k poses: eager 16.8 ms, deferred 15.4 ms;
  // a short first chunk of 1024 poses to shrink the exposed part was measured too and gave the gain back: 16.8 ms.)
  const bool lazy = h->lazy.n == n && h->lazy.upto < n;
  const int first = chunk;
  for (int p0 = 0, nb_next = first; p0 < n; p0 += nb_next, nb_next = chunk)
  for (int r = 0; r < R; r++) {
    const int nb = std::min(nb_next, n - p0);
    const float* rot = r > 0 ? h->d_rot.p + ((size_t)r * n + p0) * 9 : nullptr;
    if (lazy && r == 0) {
      // copies on the auxiliary stream (they do not queue behind the previous chunk's kernels); both streams wait for them
      stage_range(h, p0 + nb, h->aux);
      GB_CUDA(cudaEventRecord(h->ev_staged, h->aux));
      GB_CUDA(cudaStreamWaitEvent(h->stream, h->ev_staged, 0));
    }
    for (auto& Gp : h->groups) {
      GridGroup& G = *Gp;
      if (h->precision == GB_PRECISION_FP32) {
        voxelize_chunk_f32(h, G, p0, nb, rot);
        for (int mi : G.model_idx) {
          const Model& Mo = *h->models[mi];
          const size_t slot = (size_t)mi * R + r;
          h->launches += forward_fp32(Mo, G.grid.p, nb, h->ws32, h->d_out3.p, h->stream, &h->prof);
          launch_head_post(h->d_out3.p, nb, Mo.skip_softmax, Mo.apply_logistic_loss, h->d_pose.p + slot * n + p0,
                           h->d_aff.p + slot * n + p0, h->d_loss.p + slot * n + p0, h->stream, Mo.arch == GB_ARCH_OVERLAP);
          h->launches++;
        }
      } else {
        TcPoseBatch pb{G.rec_xyzr.p, G.rec_ch.p, G.n_rec, G.lig_xyzr.p, G.lig_ch.p, G.lig_off.p + p0,
                       h->d_centers.p + 3 * (size_t)p0, nb, G.max_pose_atoms, G.n_channels, G.rec.n_channels,
                       G.sig.resolution, G.sig.dimension, rot};
        TcGridWorkspace& gw = G.tc_grid;
        const int buf = (int)(gw.iter++ & 1u);
        // aux stream: wait until the network has finished reading this buffer two chunks ago, then voxelise into it
        cudaStream_t vs = h->overlap ? h->aux : h->stream;
        if (gw.consumed_valid[buf]) GB_CUDA(cudaStreamWaitEvent(vs, gw.consumed[buf], 0));
        // overlap: the voxeliser of this chunk becomes runnable only once the network of the previous chunk has
        // been handed to the GPU, so the (earlier-launched) conv CTAs keep their 2 slots per SM and voxeliser CTAs
        // fill the shared memory / thread slots that are left
        if (h->overlap && gw.started_valid) GB_CUDA(cudaStreamWaitEvent(vs, gw.started[buf ^ 1], 0));
        int kinds = 0;
        for (int mi : G.model_idx) kinds |= 1 << tc_grid_kind(*h->models[mi], false);
        h->launches += tc_prepare_grid(pb, gw, buf, kinds, vs, &h->prof);
        GB_CUDA(cudaEventRecord(gw.ready[buf], vs));
        GB_CUDA(cudaStreamWaitEvent(h->stream, gw.ready[buf], 0));
        if (h->overlap) { GB_CUDA(cudaEventRecord(gw.started[buf], h->stream)); gw.started_valid = true; }
        for (size_t k = 0; k < G.model_idx.size(); k++) {
          const int mi = G.model_idx[k];
          const Model& Mo = *h->models[mi];
          const size_t slot = (size_t)mi * R + r;
          h->launches += tc_forward(Mo, pb, gw.x0[tc_grid_kind(Mo, false)][buf], h->ws_tc, h->d_out3.p, h->stream, &h->prof,
                                    k + 1 == G.model_idx.size() ? gw.consumed[buf] : nullptr);
          launch_head_post(h->d_out3.p, nb, Mo.skip_softmax, Mo.apply_logistic_loss, h->d_pose.p + slot * n + p0,
                           h->d_aff.p + slot * n + p0, h->d_loss.p + slot * n + p0, h->stream, Mo.arch == GB_ARCH_OVERLAP);
          h->launches++;
        }
        gw.consumed_valid[buf] = true;
      }
    }
  }
  launch_ensemble(h->d_pose.p, h->d_aff.p, h->d_loss.p, M, n, n, h->d_final.p, h->d_final.p + n, h->d_final.p + 2 * (size_t)n,
                  h->d_final.p + 3 * (size_t)n, h->stream);
  h->launches++;
  GB_CUDA(cudaGetLastError());
  GB_API_END
}

int gb_cnn_fetch(gb_cnn* h, float* score, float* affinity, float* loss, float* variance) {
  GB_API_BEGIN
  GB_CHECK(h, "null handle");
  GB_CUDA(cudaSetDevice(h->device));
  const int n = h->n_staged;
  if (n == 0) return GB_OK;
  h->h_final.ensure(4 * (size_t)n);
  GB_CUDA(cudaMemcpyAsync(h->h_final.p, h->d_final.p, 4 * (size_t)n * sizeof(float), cudaMemcpyDeviceToHost, h->stream));
  GB_CUDA(cudaStreamSynchronize(h->stream));
  float* dst[4] = {score, affinity, loss, variance};
  for (int q = 0; q < 4; q++)
    if (dst[q]) memcpy(dst[q], h->h_final.p + (size_t)q * n, (size_t)n * sizeof(float));
  GB_API_END
}

int gb_cnn_fetch_device(gb_cnn* h, float* device_dst) {
  GB_API_BEGIN
  GB_CHECK(h && device_dst, "null argument");
  GB_CUDA(cudaSetDevice(h->device));
  const int n = h->n_staged;
  if (n == 0) return GB_OK;
  GB_CUDA(cudaMemcpyAsync(device_dst, h->d_final.p, 4 * (size_t)n * sizeof(float), cudaMemcpyDeviceToDevice, h->stream));
  GB_CUDA(cudaStreamSynchronize(h->stream));
  GB_API_END
}

int gb_cnn_score_batch(gb_cnn* h, const float* lig_xyz, const int32_t* lig_type, const int32_t* pose_offsets,
                       int n_poses, const float* centers, float* score, float* affinity, float* loss,
                       float* variance) {
  // staging is deferred into gb_cnn_run_staged (chunk by chunk, overlapped with the device work)
  try {
    stage_prepare(h, lig_xyz, lig_type, pose_offsets, n_poses, centers);
  } catch (const gb::Error& e) {
    gb::set_last_error(e.what());
    return e.code;
  } catch (const std::exception& e) {
    gb::set_last_error(e.what());
    return GB_ERR_INTERNAL;
  }
  int rc = gb_cnn_run_staged(h);
  if (h) h->lazy = gb_cnn::LazyInput();
  if (rc) return rc;
  return gb_cnn_fetch(h, score, affinity, loss, variance);
}

int gb_cnn_score_batch_models(gb_cnn* h, const float* lig_xyz, const int32_t* lig_type, const int32_t* pose_offsets,
                              int n_poses, const float* centers, float* pose, float* affinity, float* loss) {
  int rc = gb_cnn_stage_poses(h, lig_xyz, lig_type, pose_offsets, n_poses, centers);
  if (rc) return rc;
  rc = gb_cnn_run_staged(h);
  if (rc) return rc;
  GB_API_BEGIN
  const size_t cnt = (size_t)h->models.size() * n_rotations(h) * n_poses;  // slots model * R + rotation
  if (cnt) {
    if (pose) GB_CUDA(cudaMemcpyAsync(pose, h->d_pose.p, cnt * sizeof(float), cudaMemcpyDeviceToHost, h->stream));
    if (affinity) GB_CUDA(cudaMemcpyAsync(affinity, h->d_aff.p, cnt * sizeof(float), cudaMemcpyDeviceToHost, h->stream));
    if (loss) GB_CUDA(cudaMemcpyAsync(loss, h->d_loss.p, cnt * sizeof(float), cudaMemcpyDeviceToHost, h->stream));
  }
  GB_CUDA(cudaStreamSynchronize(h->stream));
  GB_API_END
}

int gb_cnn_profile_read(gb_cnn* h, int index, char* name, int name_cap, double* total_ms, int64_t* count) {
  GB_API_BEGIN
  GB_CHECK(h, "null handle");
  GB_CUDA(cudaSetDevice(h->device));
  GB_CUDA(cudaStreamSynchronize(h->stream));
  h->prof.resolve();
  if (index < 0 || index >= (int)h->prof.names.size()) return 1;
  if (name && name_cap > 0) { strncpy(name, h->prof.names[index].c_str(), name_cap - 1); name[name_cap - 1] = 0; }
  if (total_ms) *total_ms = h->prof.total_ms[index];
  if (count) *count = h->prof.count[index];
  GB_API_END
}
int gb_cnn_profile_reset(gb_cnn* h) {
  GB_API_BEGIN
  GB_CHECK(h, "null handle");
  GB_CUDA(cudaSetDevice(h->device));
  GB_CUDA(cudaStreamSynchronize(h->stream));
  h->prof.reset();
  GB_API_END
}

int gb_cnn_debug_read(gb_cnn* h, const char* name, void* out, size_t cap_bytes, size_t* nbytes) {
  GB_API_BEGIN
  GB_CHECK(h && name && nbytes, "null argument");
  static const char* const names[] = {"x0", "y3", "x2", "x4", "y5", "b0", "b1", "b2"};
  int idx = -1;
  for (int i = 0; i < 8; i++)
    if (std::string(name) == names[i]) idx = i;
  if (idx < 0) throw Error(GB_ERR_USAGE, "unknown debug buffer");
  size_t bytes = 0;
  const void* src = tc_debug_buffer(idx, &bytes);
  GB_CHECK(src != nullptr, "no fast-path pass has run on this thread");
  *nbytes = bytes;
  if (out) {
    GB_CHECK(cap_bytes >= bytes, "debug buffer too small");
    GB_CUDA(cudaSetDevice(h->device));
    GB_CUDA(cudaStreamSynchronize(h->stream));
    GB_CUDA(cudaMemcpy(out, src, bytes, cudaMemcpyDeviceToHost));
  }
  GB_API_END
}

// TorchModel::forward(compute_gradient = true) + CNNTorchScorer::score's accumulation (cnn_torch_scorer.cpp:164-179):
// fp32 forward keeping activations, backward of the CE loss to the grid, GridMaker::backward to the ligand atoms,
// mean over the models of the ensemble.
int gb_cnn_score_grad(gb_cnn* h, const float* lig_xyz, const int32_t* lig_type, const int32_t* pose_offsets, int n_poses,
                      const float* centers, float* score, float* affinity, float* loss, float* variance,
                      float* dlig_xyz, float* drec_xyz) {
  int rc = gb_cnn_stage_poses(h, lig_xyz, lig_type, pose_offsets, n_poses, centers);
  if (rc) return rc;
  GB_API_BEGIN
  GB_CHECK(dlig_xyz, "dlig_xyz must not be NULL");
  // getReceptorGradient (torch_model.cpp:226-232): the reference scores ONE pose per call and reads the gradient of its
  // (flexible) receptor atoms afterwards; the batch form keeps that contract -- with a shared receptor a per-pose
  // receptor gradient only exists for a single pose
  if (drec_xyz && h->n_staged != 1)
    throw Error(GB_ERR_USAGE, "receptor gradients (drec_xyz) are defined for a single pose per call");
  const int n_rec_in = (int)h->rec_type.size();
  if (drec_xyz)
    for (int i = 0; i < 3 * n_rec_in; i++) drec_xyz[i] = 0.f;
  GB_CUDA(cudaSetDevice(h->device));
  // M = ensemble members = models x rotations (cnn_torch_scorer.cpp:127-179: outputs and gradients are accumulated
  // over both loops and divided by cnt); a rotated evaluation's gradient is rotated back by the atom-gradient kernels
  const int n = h->n_staged, R = n_rotations(h), M = (int)h->models.size() * R;
  const int n_in = h->n_input_atoms;
  for (int i = 0; i < 3 * n_in; i++) dlig_xyz[i] = 0.f;
  if (n == 0) return GB_OK;
  upload_rotations(h);
  bool all_default2018 = true;
  for (Model* m : h->models) {
    if (m->arch != GB_ARCH_DEFAULT2018 && m->arch != GB_ARCH_DENSE && m->arch != GB_ARCH_OVERLAP)
      throw Error(GB_ERR_USAGE, "gradient path is implemented for the default2018 and dense families only (model " + m->name + ")");
    all_default2018 &= m->arch == GB_ARCH_DEFAULT2018;
  }
  h->d_pose.ensure((size_t)M * n);
  h->d_aff.ensure((size_t)M * n);
  h->d_loss.ensure((size_t)M * n);
  h->d_final.ensure(4 * (size_t)n);
  // precision "fp16": forward AND backward on the tensor-core path (gb_cnn_tc_grad.cu); "fp32": validation kernels
  // the tensor-core backward covers the default2018 family; an ensemble with a dense member runs the fp32 kernels
  const bool fast = h->precision != GB_PRECISION_FP32 && all_default2018;
  const int chunk = fast ? (h->max_batch > 0 ? h->max_batch : 1024) : (h->max_batch > 0 ? std::min(h->max_batch, 8) : 8);
  h->d_out3.ensure(3 * (size_t)std::max(chunk, chunk_size(h)));
  for (auto& Gp : h->groups) {
    GridGroup& G = *Gp;
    G.lig_grad.ensure(3 * (size_t)std::max(G.n_staged_atoms, 1));
    GB_CUDA(cudaMemsetAsync(G.lig_grad.p, 0, 3 * (size_t)std::max(G.n_staged_atoms, 1) * sizeof(float), h->stream));
  }
  // grow-only members, not locals: this is the per-BFGS-step call of CNN refinement, and a cudaFree per call is a
  // device-wide synchronisation that stalls every other handle's stream
  DevBuf<float>& tmp_grad = h->tmp_grad;
  DevBuf<float>& tmp_rec = h->tmp_rec;
  if (drec_xyz)
    for (auto& Gp : h->groups) {
      GridGroup& G = *Gp;
      G.rec_grad.ensure(3 * (size_t)std::max(G.n_rec, 1));
      GB_CUDA(cudaMemsetAsync(G.rec_grad.p, 0, 3 * (size_t)std::max(G.n_rec, 1) * sizeof(float), h->stream));
    }
  for (int p0 = 0; p0 < n; p0 += chunk)
  for (int r = 0; r < R; r++) {
    const int nb = std::min(chunk, n - p0);
    const float* rot = r > 0 ? h->d_rot.p + ((size_t)r * n + p0) * 9 : nullptr;
    for (auto& Gp : h->groups) {
      GridGroup& G = *Gp;
      tmp_grad.ensure(3 * (size_t)std::max(G.n_staged_atoms, 1));
      const int npts = (int)std::lround(G.sig.dimension / G.sig.resolution) + 1;
      TcPoseBatch pb{G.rec_xyzr.p, G.rec_ch.p, G.n_rec, G.lig_xyzr.p, G.lig_ch.p, G.lig_off.p + p0,
                     h->d_centers.p + 3 * (size_t)p0, nb, G.max_pose_atoms, G.n_channels, G.rec.n_channels,
                     G.sig.resolution, G.sig.dimension, rot};
      if (fast) {
        TcGridWorkspace& gw = G.tc_grid;
        h->launches += tc_prepare_grid(pb, gw, 0, 1 << 0, h->stream, &h->prof);
        gw.consumed_valid[0] = gw.consumed_valid[1] = false;  // everything here is ordered on the main stream
      } else {
        voxelize_chunk_f32(h, G, p0, nb, rot);
        h->d_dgrid.ensure((size_t)nb * G.n_channels * npts * npts * npts);
      }
      for (int mi : G.model_idx) {
        const Model& Mo = *h->models[mi];
        // the overlay test model IS a logistic-loss model (loss = -log out[1], no softmax): its backward kernel handles that
        if ((Mo.apply_logistic_loss || Mo.skip_softmax) && Mo.arch != GB_ARCH_OVERLAP)
          throw Error(GB_ERR_USAGE, "gradient of apply_logistic_loss / skip_softmax models is not implemented");
        if (fast) {
          h->launches += tc_forward(Mo, pb, G.tc_grid.x0[0][0], h->ws_tc, h->d_out3.p, h->stream, &h->prof, nullptr, true);
        } else {
          h->launches += forward_backward_fp32(Mo, G.grid.p, nb, h->ws_grad, h->d_out3.p, h->d_dgrid.p, h->stream, &h->prof);
        }
        const size_t slot = (size_t)mi * R + r;
        launch_head_post(h->d_out3.p, nb, Mo.skip_softmax, Mo.apply_logistic_loss, h->d_pose.p + slot * n + p0,
                         h->d_aff.p + slot * n + p0, h->d_loss.p + slot * n + p0, h->stream, Mo.arch == GB_ARCH_OVERLAP);
        // ligand atoms of this chunk: accumulate (scaled 1/M) into the group's gradient array
        const bool want_rec = drec_xyz && G.n_rec > 0;
        if (want_rec) tmp_rec.ensure(3 * (size_t)G.n_rec);
        if (fast) {
          h->launches += tc_backward(Mo, pb, h->ws_tc, h->d_out3.p, tmp_grad.p, h->stream, &h->prof,
                                     want_rec ? G.rec_off.p : nullptr, want_rec ? tmp_rec.p : nullptr);
        } else {
          launch_grid_backward(G.lig_xyzr.p, G.lig_ch.p, G.lig_off.p + p0, G.max_pose_atoms, h->d_centers.p + 3 * (size_t)p0, nb,
                               G.n_channels, npts, G.sig.resolution, G.sig.dimension, h->d_dgrid.p, tmp_grad.p, h->stream, rot);
          h->launches++;
          if (want_rec) {  // the receptor as the single pose's second atom set
            launch_grid_backward(G.rec_xyzr.p, G.rec_ch.p, G.rec_off.p, G.n_rec, h->d_centers.p, 1, G.n_channels, npts,
                                 G.sig.resolution, G.sig.dimension, h->d_dgrid.p, tmp_rec.p, h->stream, rot);
            h->launches++;
          }
        }
        if (want_rec) {
          launch_axpy_range(tmp_rec.p, G.rec_grad.p, 0, 3 * G.n_rec, 1.0f / (float)M, h->stream);
          h->launches++;
        }
        launch_axpy_range(tmp_grad.p, G.lig_grad.p, G.h_lig_off.p[p0] * 3, G.h_lig_off.p[p0 + nb] * 3, 1.0f / (float)M,
                          h->stream);
        h->launches += 2;
      }
    }
  }
  launch_ensemble(h->d_pose.p, h->d_aff.p, h->d_loss.p, M, n, n, h->d_final.p, h->d_final.p + n, h->d_final.p + 2 * (size_t)n,
                  h->d_final.p + 3 * (size_t)n, h->stream);
  h->launches++;
  // typed maps may differ between grid groups: every group contributes the gradient of ITS models (already /M)
  for (auto& Gp : h->groups) {
    GridGroup& G = *Gp;
    if (!G.n_staged_atoms) continue;
    G.h_lig_grad.ensure(3 * (size_t)G.n_staged_atoms);
    GB_CUDA(cudaMemcpyAsync(G.h_lig_grad.p, G.lig_grad.p, 3 * (size_t)G.n_staged_atoms * sizeof(float), cudaMemcpyDeviceToHost,
                            h->stream));
  }
  if (drec_xyz)
    for (auto& Gp : h->groups) {
      GridGroup& G = *Gp;
      if (!G.n_rec) continue;
      G.h_rec_grad.ensure(3 * (size_t)G.n_rec);
      GB_CUDA(cudaMemcpyAsync(G.h_rec_grad.p, G.rec_grad.p, 3 * (size_t)G.n_rec * sizeof(float), cudaMemcpyDeviceToHost, h->stream));
    }
  GB_CUDA(cudaStreamSynchronize(h->stream));
  if (drec_xyz)
    for (auto& Gp : h->groups) {
      GridGroup& G = *Gp;
      for (int a = 0; a < G.n_rec; a++)
        for (int d = 0; d < 3; d++) drec_xyz[3 * G.h_rec_src[a] + d] += G.h_rec_grad.p[3 * a + d];
    }
  for (auto& Gp : h->groups) {
    GridGroup& G = *Gp;
    for (int a = 0; a < G.n_staged_atoms; a++) {
      const int src = G.h_lig_src.p[a];
      for (int d = 0; d < 3; d++) dlig_xyz[3 * src + d] += G.h_lig_grad.p[3 * a + d];
    }
  }
  GB_CUDA(cudaGetLastError());
  rc = gb_cnn_fetch(h, score, affinity, loss, variance);
  if (rc) return rc;
  GB_API_END
}

void* gb_cnn_stream(gb_cnn* h) { return h ? (void*)h->stream : nullptr; }
int64_t gb_cnn_kernel_launches(gb_cnn* h) { return h ? h->launches : 0; }

int gb_cnn_voxelize(gb_cnn* h, int model_index, const float* lig_xyz, const int32_t* lig_type,
                    const int32_t* pose_offsets, int n_poses, const float* centers, float* grid_out) {
  int rc = gb_cnn_stage_poses(h, lig_xyz, lig_type, pose_offsets, n_poses, centers);
  if (rc) return rc;
  GB_API_BEGIN
  GB_CHECK(model_index >= 0 && model_index < (int)h->models.size() && grid_out, "bad voxelize arguments");
  GridGroup& G = *h->groups[h->model_group[model_index]];
  const int npts = h->models[model_index]->npts;
  const size_t per = (size_t)G.n_channels * npts * npts * npts;
  const int chunk = 8;
  for (int p0 = 0; p0 < n_poses; p0 += chunk) {
    const int nb = std::min(chunk, n_poses - p0);
    voxelize_chunk_f32(h, G, p0, nb);
    GB_CUDA(cudaMemcpyAsync(grid_out + per * p0, G.grid.p, per * nb * sizeof(float), cudaMemcpyDeviceToHost, h->stream));
    GB_CUDA(cudaStreamSynchronize(h->stream));
  }
  GB_API_END
}

}  // extern "C"
