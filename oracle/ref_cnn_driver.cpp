// oracle/ref_cnn_driver.cpp -- C-ABI driver over the REFERENCE's own CNN scoring host code, compiled where it lies under /root/reference:
// lib/torch_model.cpp (TorchModel: metadata, type maps, make_coordset, centre, rec+lig merge, head post-processing, autograd backward
// and gradient split) and lib/cnn_torch_scorer.cpp (CNNTorchScorer: built-in model names, ensemble mean / variance, gradient accumulation
// into the model) on top of lib/dl_scorer.cpp (setLigand / setReceptor, already in libgnina_vina_ref.so).  The networks are the
// reference's own TorchScript files run by libtorch on the CPU.  libmolgrid -- third party, absent -- is the one stand-in with
// arithmetic: oracle/ref_shim/libmolgrid forwards to oracle/gridmaker_ref.c.  TEST INFRASTRUCTURE (oracle/): pins oracle/pipeline.py and
// generates tests/golden/cnn_ref_kat.npz; built only where /root/reference exists.
#include <dirent.h>
#include <cstdio>
#include <cstring>
#include <fstream>
#include <memory>
#include <string>
#include <vector>

#include "cnn_torch_scorer.h"
#include "torch_models.h"
#include <boost/iostreams/device/array.hpp>
#include <boost/iostreams/stream.hpp>

// what make_model_cpp.py generates at the reference's build time from lib/models/*.pt (linker-embedded there, read from the
// reference tree here)
boost::unordered_map<std::string, std::pair<char*, char*> > torch_models;
static std::vector<std::unique_ptr<std::vector<char>>> g_blobs;
static std::string g_err;

extern "C" void* gref_model_ptr(void* p);   // libgnina_vina_ref.so: the `model` inside a gref_model_create handle

namespace {
template <class F> int guarded(F&& f) {
  try { f(); return 0; }
  catch (const std::exception& e) { g_err = e.what(); }
  catch (const internal_error& e) { g_err = "internal_error " + e.file + ":" + std::to_string(e.line); }
  catch (...) { g_err = "unknown exception"; }
  return 1;
}
struct RefCNN { std::unique_ptr<DLScorer> s; };
}  // namespace

extern "C" {

const char* gcref_last_error() { return g_err.c_str(); }

// name = file stem with '.' -> '_' (make_model_cpp.py:27-29)
int gcref_load_models(const char* dir) {
  DIR* d = opendir(dir);
  if (!d) return -1;
  int n = 0;
  while (dirent* e = readdir(d)) {
    std::string f = e->d_name;
    if (f.size() < 4 || f.substr(f.size() - 3) != ".pt") continue;
    std::ifstream in(std::string(dir) + "/" + f, std::ios::binary);
    std::unique_ptr<std::vector<char>> b(new std::vector<char>((std::istreambuf_iterator<char>(in)), std::istreambuf_iterator<char>()));
    std::string name = f.substr(0, f.size() - 3);
    for (char& c : name) if (c == '.') c = '_';
    torch_models[name] = std::make_pair(b->data(), b->data() + b->size());
    g_blobs.push_back(std::move(b));
    n++;
  }
  closedir(d);
  return n;
}

// names: built-in model names (0 names and 0 files = the reference's default ensemble); files: --cnn_models paths
void* gcref_scorer_create(const char** names, int n_names, const char** files, int n_files, unsigned rotations, unsigned seed,
                          const float* cnn_center) {
  RefCNN* R = new RefCNN;
  int rc = guarded([&] {
    cnn_options o;
    for (int i = 0; i < n_names; i++) o.cnn_model_names.push_back(names[i]);
    for (int i = 0; i < n_files; i++) o.cnn_models.push_back(files[i]);
    o.cnn_rotations = rotations; o.seed = seed; o.cnn_scoring = CNNall;
    if (cnn_center) o.cnn_center = vec(cnn_center[0], cnn_center[1], cnn_center[2]);
    R->s.reset(new CNNTorchScorer<false>(o, nullptr));
  });
  if (rc) { delete R; return nullptr; }
  return R;
}
void gcref_scorer_destroy(void* p) { delete (RefCNN*)p; }
// as a DLScorer*, for libgnina_vina_ref.so's gref_minimize_dl / gref_refine_dl (quasi_newton + non_cache_cnn over any DLScorer)
void* gcref_scorer_dl(void* p) { return ((RefCNN*)p)->s.get(); }

// the model names the constructor resolved (aliases and ensembles expanded) are private; the expansion is observable through
// the number of evaluations only, so the driver reports what the public interface offers
float gcref_grid_dim(void* p) { auto* t = dynamic_cast<CNNTorchScorer<false>*>(((RefCNN*)p)->s.get()); return t ? t->get_grid_dim() : 0.f; }
float gcref_grid_res(void* p) { auto* t = dynamic_cast<CNNTorchScorer<false>*>(((RefCNN*)p)->s.get()); return t ? t->get_grid_res() : 0.f; }

// CNNTorchScorer::score(m, compute_gradient, affinity, loss, variance) on the coordinates the model holds
// out4 = score, affinity, loss, variance; minus_forces [n_movable][3] = what the call left in the model
int gcref_score(void* p, void* model_handle, int compute_gradient, float* out4, float* minus_forces) {
  return guarded([&] {
    model& m = *(model*)gref_model_ptr(model_handle);
    float aff = 0, loss = 0, var = 0;
    const float s = ((RefCNN*)p)->s->score(m, compute_gradient != 0, aff, loss, var);
    out4[0] = s; out4[1] = aff; out4[2] = loss; out4[3] = var;
    if (minus_forces)
      for (sz i = 0; i < m.minus_forces.size(); i++) for (int k = 0; k < 3; k++) minus_forces[3 * i + k] = m.minus_forces[i][k];
  });
}

// DLScorer::set_center_from_model + get_center + CNNTorchScorer::set_bounding_box -> centre, box begin / end / n
int gcref_center_and_box(void* p, void* model_handle, float* center, float* begin, float* end, int* n) {
  return guarded([&] {
    model& m = *(model*)gref_model_ptr(model_handle);
    DLScorer& s = *((RefCNN*)p)->s;
    s.set_center_from_model(m);
    const vec c = s.get_center();
    grid_dims gd;
    s.set_bounding_box(gd);
    for (int i = 0; i < 3; i++) { center[i] = c[i]; begin[i] = gd[i].begin; end[i] = gd[i].end; n[i] = (int)gd[i].n; }
  });
}

}  // extern "C"

// ---- the integration adapter EXECUTED on the CPU -----------------------------------------------------------------------------------
// integration/cnn_b200_scorer.h (CNNB200Scorer : DLScorer, the one class a gnina maintainer adds) calls twelve entry points of
// include/gnina_b200.h.  Here those twelve names are redirected to a stand-in that honours the C ABI's CONTRACT with the reference's
// own TorchModel as the network (one pose per call; ensemble arithmetic of CNNTorchScorer::score; the by-atom gradient averaged over
// the models) -- so the adapter's own code (setLigand / setReceptor reuse, receptor upload, centre option, gradient scatter,
// add_minus_forces, fresh_copy, set_bounding_box, name resolution against the packaged blobs) runs inside the reference's
// non_cache_cnn / quasi_newton and can be compared with CNNTorchScorer in the same process.  That the device library honours the same
// contract is what the GPU parity tests check.
#include "gnina_b200.h"
struct gb_model { std::shared_ptr<TorchModel<false>> tm; std::string name; int refs = 1; };
struct gb_cnn {
  std::vector<gb_model*> models;
  std::vector<float3> rec; std::vector<smt> rec_t;
};
static std::string g_mock_err;
extern "C" {
const char* mockgb_last_error(void) { return g_mock_err.c_str(); }
int mockgb_model_load(const char* path, int, gb_model** out) {
  std::string stem = path;
  const size_t slash = stem.rfind('/');
  if (slash != std::string::npos) stem = stem.substr(slash + 1);
  const size_t dot = stem.rfind('.');
  const std::string ext = dot == std::string::npos ? "" : stem.substr(dot);
  stem = stem.substr(0, dot);
  try {
    gb_model* m = new gb_model; m->name = stem;
    if (ext == ".gbw" && torch_models.count(stem)) {   // a packaged blob <-> the TorchScript file it was converted from
      boost::iostreams::basic_array_source<char> src(torch_models[stem].first, torch_models[stem].second - torch_models[stem].first);
      boost::iostreams::stream<boost::iostreams::basic_array_source<char>> in(src);
      m->tm = std::make_shared<TorchModel<false>>(in, stem, nullptr);
    } else {
      std::ifstream in(path, std::ios::binary);
      if (!in) { delete m; g_mock_err = std::string("Could not read torch model ") + path; return GB_ERR_USAGE; }
      m->tm = std::make_shared<TorchModel<false>>(in, stem, nullptr);
    }
    *out = m; return GB_OK;
  } catch (...) { g_mock_err = std::string("Could not read torch model ") + path; return GB_ERR_USAGE; }
}
void mockgb_model_release(gb_model* m) { if (m && --m->refs == 0) delete m; }
int mockgb_model_get_info(const gb_model* m, gb_model_info* info) {
  std::memset(info, 0, sizeof *info);
  info->dimension = m->tm->get_grid_dim(); info->resolution = m->tm->get_grid_res();
  std::strncpy(info->name, m->name.c_str(), sizeof info->name - 1);
  return GB_OK;
}
int mockgb_cnn_create(gb_model* const* models, int n, int, gb_cnn** out) {
  gb_cnn* h = new gb_cnn;
  for (int i = 0; i < n; i++) { models[i]->refs++; h->models.push_back(models[i]); }   // handles keep their models alive
  *out = h; return GB_OK;
}
int mockgb_cnn_clone(const gb_cnn* h, gb_cnn** out) {
  gb_cnn* c = new gb_cnn(*h);
  for (gb_model* m : c->models) m->refs++;
  *out = c; return GB_OK;
}
void mockgb_cnn_destroy(gb_cnn* h) { for (gb_model* m : h->models) mockgb_model_release(m); delete h; }
int mockgb_cnn_num_models(const gb_cnn* h) { return (int)h->models.size(); }
int mockgb_cnn_set_option(gb_cnn*, const char* key, double v) {
  if (std::string(key) == "cnn_rotation" && v > 1) { g_mock_err = "rotations are not restated"; return GB_ERR_USAGE; }
  return GB_OK;
}
int mockgb_cnn_set_receptor(gb_cnn* h, const float* xyz, const int32_t* t, int n) {
  h->rec.resize(n); h->rec_t.resize(n);
  for (int i = 0; i < n; i++) { h->rec[i] = make_float3(xyz[3 * i], xyz[3 * i + 1], xyz[3 * i + 2]); h->rec_t[i] = (smt)t[i]; }
  return GB_OK;
}
static int mock_score_one(gb_cnn* h, const float* xyz, const int32_t* t, int n, const float* centers, float* score, float* affinity,
                          float* loss, float* variance, float* grad) {
  std::vector<float3> lig(n); std::vector<smt> lt(n);
  for (int i = 0; i < n; i++) { lig[i] = make_float3(xyz[3 * i], xyz[3 * i + 1], xyz[3 * i + 2]); lt[i] = (smt)t[i]; }
  const vec center = centers ? vec(centers[0], centers[1], centers[2]) : vec(NAN, NAN, NAN);
  double s = 0; float a = 0, l = 0; std::vector<float> affs; std::vector<gfloat3> g;
  if (grad) std::fill(grad, grad + 3 * (size_t)n, 0.f);
  const unsigned cnt = (unsigned)h->models.size();
  for (gb_model* m : h->models) {
    const std::vector<float> o = m->tm->forward(h->rec, h->rec_t, lig, lt, center, false, grad != nullptr);
    s += o[0]; a += o[1]; l += o[2]; affs.push_back(o[1]);
    if (grad) { m->tm->getLigandGradient(g); for (int i = 0; i < n; i++) { grad[3 * i] += g[i].x; grad[3 * i + 1] += g[i].y; grad[3 * i + 2] += g[i].z; } }
  }
  if (grad && cnt > 1) { const float sc = 1.0 / cnt; for (size_t i = 0; i < 3 * (size_t)n; i++) grad[i] *= sc; }
  a /= cnt; l /= cnt; s /= cnt;
  float var = 0;
  if (affs.size() > 1) { float sum = 0; for (float q : affs) { float d = a - q; d *= d; sum += d; } var = sum / affs.size(); }
  if (score) *score = (float)s;
  if (affinity) *affinity = a;
  if (loss) *loss = l;
  if (variance) *variance = var;
  return GB_OK;
}
// poses are independent: one after the other; every output array is nullable as in the C ABI
static int mock_score(gb_cnn* h, const float* xyz, const int32_t* t, const int32_t* offs, int n_poses, const float* centers, float* score,
                      float* affinity, float* loss, float* variance, float* grad) {
  for (int p = 0; p < n_poses; p++) {
    const int a = offs[p], n = offs[p + 1] - offs[p];
    const int rc = mock_score_one(h, xyz + 3 * a, t + a, n, centers ? centers + 3 * p : nullptr, score ? score + p : nullptr,
                                  affinity ? affinity + p : nullptr, loss ? loss + p : nullptr, variance ? variance + p : nullptr,
                                  grad ? grad + 3 * a : nullptr);
    if (rc != GB_OK) return rc;
  }
  return GB_OK;
}
int mockgb_cnn_score_batch(gb_cnn* h, const float* xyz, const int32_t* t, const int32_t* offs, int n_poses, const float* centers, float* score,
                           float* affinity, float* loss, float* variance) {
  try { return mock_score(h, xyz, t, offs, n_poses, centers, score, affinity, loss, variance, nullptr); }
  catch (const std::exception& e) { g_mock_err = e.what(); return GB_ERR_INTERNAL; }
}
int mockgb_cnn_score_grad(gb_cnn* h, const float* xyz, const int32_t* t, const int32_t* offs, int n_poses, const float* centers, float* score,
                          float* affinity, float* loss, float* variance, float* grad, float* rec_grad) {
  if (rec_grad) { g_mock_err = "receptor gradients (flexible residues) are out of scope"; return GB_ERR_USAGE; }
  try { return mock_score(h, xyz, t, offs, n_poses, centers, score, affinity, loss, variance, grad); }
  catch (const std::exception& e) { g_mock_err = e.what(); return GB_ERR_INTERNAL; }
}
}
extern "C" int mockgb_initialize_cuda(int) { return 0; }
#define gb_initialize_cuda mockgb_initialize_cuda
#define gb_last_error mockgb_last_error
#define gb_model_load mockgb_model_load
#define gb_model_release mockgb_model_release
#define gb_model_get_info mockgb_model_get_info
#define gb_cnn_create mockgb_cnn_create
#define gb_cnn_clone mockgb_cnn_clone
#define gb_cnn_destroy mockgb_cnn_destroy
#define gb_cnn_num_models mockgb_cnn_num_models
#define gb_cnn_set_option mockgb_cnn_set_option
#define gb_cnn_set_receptor mockgb_cnn_set_receptor
#define gb_cnn_score_batch mockgb_cnn_score_batch
#define gb_cnn_score_grad mockgb_cnn_score_grad
#include "cnn_b200_scorer.h"
#include "gnina_b200_minimize.hpp"   // gb::LigandTree, gb::minimize_poses, gb::CnnBatchEnergy (redirected to the stand-in like the rest)
#include "docking_b200.h"            // b200::B200Ligand: the model -> gb_ligand_topology adapter

extern "C" {
// CNNB200Scorer(cnn_options, device, blob_dir) over the stand-in; copy != 0: hand out fresh_copy() of it instead (what every
// docking thread gets)
void* gcref_adapter_create(const char** names, int n_names, const char** files, int n_files, const char* blob_dir, const float* cnn_center,
                           int copy) {
  RefCNN* R = new RefCNN;
  int rc = guarded([&] {
    cnn_options o;
    for (int i = 0; i < n_names; i++) o.cnn_model_names.push_back(names[i]);
    for (int i = 0; i < n_files; i++) o.cnn_models.push_back(files[i]);
    o.cnn_scoring = CNNall;
    if (cnn_center) o.cnn_center = vec(cnn_center[0], cnn_center[1], cnn_center[2]);
    std::unique_ptr<CNNB200Scorer> a(new CNNB200Scorer(o, 0, blob_dir));
    if (copy) {
      std::shared_ptr<DLScorer> c = a->fresh_copy();
      struct Holder : DLScorer {                      // keeps the shared_ptr; forwards the interface
        std::shared_ptr<DLScorer> c; std::unique_ptr<CNNB200Scorer> parent;
        bool initialized() const override { return c->initialized(); }
        bool has_affinity() const override { return c->has_affinity(); }
        float score(model& m, float& v) override { return c->score(m, v); }
        float score(model& m, bool g, float& a, float& l, float& v) override { return c->score(m, g, a, l, v); }
        void set_center_from_model(model& m) override { c->set_center_from_model(m); }
        vec get_center() const override { return c->get_center(); }
        void set_bounding_box(grid_dims& b) const override { c->set_bounding_box(b); }
        std::shared_ptr<DLScorer> fresh_copy() const override { return c->fresh_copy(); }
      };
      Holder* hd = new Holder; hd->c = c; hd->parent = std::move(a);
      R->s.reset(hd);
    } else {
      R->s = std::move(a);
    }
  });
  if (rc) { delete R; return nullptr; }
  return R;
}
}

// ---- the product's own C++ host classes on the CPU: gb::CNNScorer + gb::NonCacheCNN (include/gnina_b200.hpp) over the same stand-in --
// non_cache_cnn's counterpart evaluated on the pose the model holds, CNN box centred on `center` (what adjust_center gave the
// reference side): -> e, forces [n_movable][3]
extern "C" int gcref_product_noncache_cnn(const char** names, int n_names, const char* weights_dir, void* model_handle, const float* begin,
                                          const float* end, const int* n, float slope, const float* center, float v, int with_deriv,
                                          int reference_force_routing, float* e, float* forces) {
  return guarded([&] {
    model& m = *(model*)gref_model_ptr(model_handle);
    std::vector<std::string> nm;
    for (int i = 0; i < n_names; i++) nm.push_back(names[i]);
    gb::CNNScorer scorer(weights_dir, nm, 0);
    std::vector<float> rx; std::vector<int32_t> rt;
    for (const atom& a : m.get_fixed_atoms()) { rt.push_back((int32_t)a.sm); for (int k = 0; k < 3; k++) rx.push_back((float)a.coords[k]); }
    scorer.set_receptor(rx.data(), rt.data(), (int)rt.size());
    gb::GridDims gd;
    for (int i = 0; i < 3; i++) { gd[i].begin = begin[i]; gd[i].end = end[i]; gd[i].n = n[i]; }
    gb::NonCacheCNN nc(scorer, gd, center, slope);
    nc.set_reference_force_routing(reference_force_routing != 0);
    const int na = (int)m.num_movable_atoms();
    std::vector<float> xyz(3 * (size_t)na), f;
    std::vector<int32_t> t(na);
    for (int i = 0; i < na; i++) { t[i] = (int32_t)m.atoms[i].sm; for (int k = 0; k < 3; k++) xyz[3 * i + k] = (float)m.coords[i][k]; }
    *e = nc.eval(xyz.data(), t.data(), na, with_deriv ? &f : nullptr, v);
    if (with_deriv) std::copy(f.begin(), f.end(), forces);
  });
}

// config 5 through the product's C++ host code: gb::CNNScorer + gb::LigandTree + gb::CnnBatchEnergy + gb::minimize_poses
// (include/gnina_b200_minimize.hpp) on n conformations of the model's ligand at once; ONE gb_cnn_score_grad call per round.
// confs [n][7+T] in/out; the kinematics use this host's sinf / cosf / acosf like the reference build
extern "C" int gcref_product_lockstep_minimize(const char** names, int n_names, const char* weights_dir, void* model_handle, const float* begin,
                                               const float* end, float slope, float* confs, int n, int maxiters, int accurate, int early_term,
                                               int reference_force_routing, float* e_out, int* evals_out, int* rounds_out) {
  return guarded([&] {
    model& m = *(model*)gref_model_ptr(model_handle);
    std::vector<std::string> nm;
    for (int i = 0; i < n_names; i++) nm.push_back(names[i]);
    gb::CNNScorer scorer(weights_dir, nm, 0);
    std::vector<float> rx; std::vector<int32_t> rt;
    for (const atom& a : m.get_fixed_atoms()) { rt.push_back((int32_t)a.sm); for (int k = 0; k < 3; k++) rx.push_back((float)a.coords[k]); }
    scorer.set_receptor(rx.data(), rt.data(), (int)rt.size());
    b200::B200Ligand BL(m);
    gb::LigandTree tree(BL.topo);
    gb::Transcendentals saved = gb::transcendentals();
    gb::transcendentals().sin = [](float x) { return sinf(x); };
    gb::transcendentals().cos = [](float x) { return cosf(x); };
    gb::transcendentals().acos = [](float x) { return acosf(x); };
    gb::CnnBatchEnergy energy(scorer.handle(), tree, begin, end, slope, scorer.info(0).dimension);
    energy.set_reference_force_routing(reference_force_routing != 0);
    energy.set_centers(confs, n);
    gb::MinimizeParams mp; mp.maxiters = maxiters; mp.accurate_line_search = accurate != 0; mp.early_term = early_term != 0;
    std::vector<int> ev; int rounds = 0;
    std::vector<float> e = gb::minimize_poses(tree, energy, confs, n, mp, &ev, &rounds);
    gb::transcendentals() = saved;
    std::copy(e.begin(), e.end(), e_out);
    std::copy(ev.begin(), ev.end(), evals_out);
    *rounds_out = rounds;
  });
}
