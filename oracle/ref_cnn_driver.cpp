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
struct RefCNN { std::unique_ptr<CNNTorchScorer<false>> s; };
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
void* gcref_scorer_dl(void* p) { return static_cast<DLScorer*>(((RefCNN*)p)->s.get()); }

// the model names the constructor resolved (aliases and ensembles expanded) are private; the expansion is observable through
// the number of evaluations only, so the driver reports what the public interface offers
float gcref_grid_dim(void* p) { return ((RefCNN*)p)->s->get_grid_dim(); }
float gcref_grid_res(void* p) { return ((RefCNN*)p)->s->get_grid_res(); }

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
    CNNTorchScorer<false>& s = *((RefCNN*)p)->s;
    s.set_center_from_model(m);
    const vec c = s.get_center();
    grid_dims gd;
    s.set_bounding_box(gd);
    for (int i = 0; i < 3; i++) { center[i] = c[i]; begin[i] = gd[i].begin; end[i] = gd[i].end; n[i] = (int)gd[i].n; }
  });
}

}  // extern "C"
