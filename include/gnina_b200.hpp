// gnina_b200.hpp — C++ host side above the C ABI (header-only, C++17, no dependencies beyond gnina_b200.h).
//
// gnina's host code is C++; this header is the host-language mirror of the reference's scorer interface for the
// accelerated path, written so that the adapter sketched in INTEGRATION.md is a few lines:
//   gb::CNNScorer   <-> CNNTorchScorer<isCUDA>  (gninasrc/lib/cnn_torch_scorer.{h,cpp}) / DLScorer (lib/dl_scorer.h:23-66)
//   gb::NonCacheCNN <-> non_cache_cnn::eval / eval_deriv   (gninasrc/lib/non_cache_cnn.cpp:33-54,79-169), default
//                       options (no empirical mixing, no user grid): CNN loss + out-of-box penalties of the search box
//                       and of the CNN box (non_cache::check_bounds(_deriv), lib/non_cache.cpp:32-50,102-123)
//   gb::VinaScorer  <-> precalculate_linear + cache + naive_non_cache final scoring (see gnina_b200.h)
//   gb::PoseBatcher <-> the pose queue SURVEY.md 8(f)-1 asks for: gnina's ligand loop (main/main.cpp:749-771, 233-269,
//                       324-346) scores one pose per call; the batcher collects poses and hands them to the batch
//                       entry point, delivering results in submission order
//   gb::read_gninatypes / write_gninatypes <-> gninatyper's typed-atom records (gninatyper/gninatyper.cpp:30-36,72-77)
// Errors: GB_ERR_USAGE -> gb::usage_error (reference: usage_error), everything else -> gb::internal_error.
#pragma once
#include <array>
#include <cmath>
#include <cstdio>
#include <functional>
#include <memory>
#include <stdexcept>
#include <string>
#include <exception>
#include <thread>
#include <vector>
#include "gnina_b200.h"

namespace gb {

struct usage_error : std::runtime_error { using std::runtime_error::runtime_error; };
struct internal_error : std::runtime_error { using std::runtime_error::runtime_error; };
inline void check(int rc) {
  if (rc == GB_OK) return;
  if (rc == GB_ERR_USAGE) throw usage_error(gb_last_error());
  throw internal_error(gb_last_error());
}

// cnn_torch_scorer.cpp:28-62: default ensemble, "fast", "default1.0"; "X_ensemble" needs the list of built-in names
inline std::vector<std::string> expand_model_names(std::vector<std::string> names,
                                                   const std::vector<std::string>& builtin = {}) {
  for (auto& n : names)
    if (n != "default1.0")
      for (auto& c : n)
        if (c == '.') c = '_';  // make_model_cpp.py:31-32
  if (names.empty()) names = {"dense_1_3", "dense_1_3_PT_KD_3", "crossdock_default2018_KD_4"};
  else if (names.size() == 1) {
    if (names[0] == "fast") names = {"all_default_to_default_1_3_1"};
    else if (names[0] == "default1.0")
      names = {"dense", "general_default2018_3", "dense_3", "crossdock_default2018", "redock_default2018_2"};
  }
  std::vector<std::string> out;
  const std::string suffix = "_ensemble";
  for (const auto& n : names) {
    if (n.size() > suffix.size() && n.compare(n.size() - suffix.size(), suffix.size(), suffix) == 0) {
      const std::string prefix = n.substr(0, n.size() - suffix.size());
      for (const auto& b : builtin)
        if (b.compare(0, prefix.size(), prefix) == 0) out.push_back(b);
    } else out.push_back(n);
  }
  return out;
}

struct Scores { std::vector<float> score, affinity, loss, variance; };

class CNNScorer {
  gb_cnn* h_ = nullptr;
  std::vector<gb_model*> models_;
  std::vector<gb_model_info> infos_;   // per model; kept by fresh copies too (they share the handle's models, not models_)
  int device_ = 0;
  int n_rec_ = 0;
  CNNScorer() = default;

 public:
  // weights_dir holds the GNB200W1 blobs (gnina_b200/weights); names as gnina spells them
  CNNScorer(const std::string& weights_dir, const std::vector<std::string>& names, int device = 0) : device_(device) {
    if (gb_initialize_cuda(device) != 0) throw internal_error("no usable CUDA device (gnina_b200 has no CPU fallback)");
    for (const auto& n : expand_model_names(names)) {
      gb_model* m = nullptr;
      const int rc = gb_model_load((weights_dir + "/" + n + ".gbw").c_str(), device, &m);
      if (rc == GB_ERR_USAGE) throw usage_error("Invalid model name: " + n);  // cnn_torch_scorer.cpp:70-72
      check(rc);
      models_.push_back(m);
      gb_model_info inf;
      check(gb_model_get_info(m, &inf));
      infos_.push_back(inf);
    }
    check(gb_cnn_create(models_.data(), (int)models_.size(), device, &h_));
  }
  CNNScorer(const CNNScorer&) = delete;
  CNNScorer& operator=(const CNNScorer&) = delete;
  ~CNNScorer() {
    if (h_) gb_cnn_destroy(h_);
    for (auto m : models_) gb_model_release(m);
  }

  bool initialized() const { return h_ && gb_cnn_num_models(h_) > 0; }
  bool has_affinity() const { return true; }
  std::unique_ptr<CNNScorer> fresh_copy() const {  // cnn_torch_scorer.h:54
    std::unique_ptr<CNNScorer> c(new CNNScorer);
    c->device_ = device_;
    c->n_rec_ = n_rec_;
    c->infos_ = infos_;
    check(gb_cnn_clone(h_, &c->h_));
    return c;
  }
  void set_option(const char* key, double v) { check(gb_cnn_set_option(h_, key, v)); }
  gb_cnn* handle() const { return h_; }   // for the C-ABI level helpers (gnina_b200_minimize.hpp's CnnBatchEnergy)
  gb_model_info info(int i = 0) const { return infos_.at(i); }
  void set_receptor(const float* xyz, const int32_t* smina_type, int n) {
    check(gb_cnn_set_receptor(h_, xyz, smina_type, n));
    n_rec_ = n;
  }

  Scores score_batch(const float* lig_xyz, const int32_t* lig_type, const int32_t* pose_offsets, int n_poses,
                     const float* centers = nullptr) {
    Scores s;
    s.score.resize(n_poses); s.affinity.resize(n_poses); s.loss.resize(n_poses); s.variance.resize(n_poses);
    check(gb_cnn_score_batch(h_, lig_xyz, lig_type, pose_offsets, n_poses, centers, s.score.data(), s.affinity.data(),
                             s.loss.data(), s.variance.data()));
    return s;
  }
  // DLScorer::score(model&, compute_gradient, affinity, loss, variance): one pose; gradient (if requested) is the
  // derivative of the loss w.r.t. every ligand atom passed = what the reference adds to m.minus_forces
  // receptor_gradient (optional, sized 3 x receptor atoms by the caller's set_receptor): getReceptorGradient, what the
  // reference adds to the flexible-residue atoms' minus_forces
  float score(const float* lig_xyz, const int32_t* lig_type, int n_atoms, bool compute_gradient, float& affinity,
              float& loss, float& variance, std::vector<float>* gradient = nullptr, const float* center = nullptr,
              std::vector<float>* receptor_gradient = nullptr) {
    if (!initialized()) return -1.0f;  // cnn_torch_scorer.cpp:107-108
    const int32_t offs[2] = {0, n_atoms};
    float s = 0;
    if (compute_gradient) {
      std::vector<float> g(3 * (size_t)n_atoms);
      if (receptor_gradient) receptor_gradient->assign(3 * (size_t)n_rec_, 0.f);
      check(gb_cnn_score_grad(h_, lig_xyz, lig_type, offs, 1, center, &s, &affinity, &loss, &variance, g.data(),
                              receptor_gradient ? receptor_gradient->data() : nullptr));
      if (gradient) *gradient = std::move(g);
    } else {
      check(gb_cnn_score_batch(h_, lig_xyz, lig_type, offs, 1, center, &s, &affinity, &loss, &variance));
    }
    return s;
  }
};

// ---- one process, several GPUs (poses are independent units, no exchange step): one CNNScorer per device,
// the poses of a batch cut into contiguous shards, every shard scored by its own host thread on its own device, results
// concatenated in pose order.  The reference has nothing like it (one scorer, one GPU); this is the C++-side counterpart
// of the torchrun pose sharding the benchmark uses.
class MultiDeviceScorer {
  std::vector<std::unique_ptr<CNNScorer>> scorers_;

 public:
  // devices may repeat (two handles on one GPU overlap their kernels); empty = every visible device
  MultiDeviceScorer(const std::string& weights_dir, const std::vector<std::string>& names, std::vector<int> devices = {}) {
    if (devices.empty())
      for (int d = 0; d < gb_device_count(); d++) devices.push_back(d);
    if (devices.empty()) throw internal_error("no usable CUDA device (gnina_b200 has no CPU fallback)");
    for (int d : devices) scorers_.emplace_back(new CNNScorer(weights_dir, names, d));
  }
  size_t n_devices() const { return scorers_.size(); }
  CNNScorer& scorer(size_t i) { return *scorers_.at(i); }
  void set_option(const char* key, double v) { for (auto& s : scorers_) s->set_option(key, v); }
  void set_receptor(const float* xyz, const int32_t* smina_type, int n) {
    for (auto& s : scorers_) s->set_receptor(xyz, smina_type, n);   // replicated: <= 128 KB per device
  }
  // shard boundaries: pose p of n goes to device p * D / n (contiguous, sizes differ by at most one)
  static std::vector<int> shard_bounds(int n_poses, int n_dev) {
    std::vector<int> b(n_dev + 1);
    for (int d = 0; d <= n_dev; d++) b[d] = (int)((long long)n_poses * d / n_dev);
    return b;
  }
  Scores score_batch(const float* lig_xyz, const int32_t* lig_type, const int32_t* pose_offsets, int n_poses,
                     const float* centers = nullptr) {
    const int D = (int)scorers_.size();
    const std::vector<int> b = shard_bounds(n_poses, D);
    Scores out;
    out.score.resize(n_poses); out.affinity.resize(n_poses); out.loss.resize(n_poses); out.variance.resize(n_poses);
    std::vector<std::exception_ptr> err(D);
    std::vector<std::thread> th;
    for (int d = 0; d < D; d++) {
      th.emplace_back([&, d] {
        try {
          const int p0 = b[d], np = b[d + 1] - b[d];
          if (np == 0) return;
          std::vector<int32_t> off(np + 1);
          for (int i = 0; i <= np; i++) off[i] = pose_offsets[p0 + i] - pose_offsets[p0];
          const int a0 = pose_offsets[p0];
          Scores s = scorers_[d]->score_batch(lig_xyz + 3 * (size_t)a0, lig_type + a0, off.data(), np, centers ? centers + 3 * (size_t)p0 : nullptr);
          std::copy(s.score.begin(), s.score.end(), out.score.begin() + p0);
          std::copy(s.affinity.begin(), s.affinity.end(), out.affinity.begin() + p0);
          std::copy(s.loss.begin(), s.loss.end(), out.loss.begin() + p0);
          std::copy(s.variance.begin(), s.variance.end(), out.variance.begin() + p0);
        } catch (...) { err[d] = std::current_exception(); }
      });
    }
    for (auto& t : th) t.join();
    for (auto& e : err)
      if (e) std::rethrow_exception(e);
    return out;
  }
};

// ---- typed-atom records: `.gninatypes` = a flat array of {float x, y, z; int32 smina_type} (gninatyper.cpp:30-36,
// written one molecule per file at :72-77).  The smina type is exactly the int32 the ABI takes, so a virtual screen
// can feed pre-typed ligands without OpenBabel.
struct TypedAtoms {
  std::vector<float> xyz;      // [n][3]
  std::vector<int32_t> type;   // [n]
  size_t size() const { return type.size(); }
};
inline TypedAtoms read_gninatypes(const std::string& path) {
  FILE* f = std::fopen(path.c_str(), "rb");
  if (!f) throw usage_error("Could not open " + path);
  TypedAtoms a;
  struct Rec { float x, y, z; int32_t t; } r;
  static_assert(sizeof(Rec) == 16, "gninatypes record");
  size_t got;
  while ((got = std::fread(&r, 1, sizeof r, f)) == sizeof r) {
    a.xyz.push_back(r.x); a.xyz.push_back(r.y); a.xyz.push_back(r.z);
    a.type.push_back(r.t);
  }
  std::fclose(f);
  if (got != 0) throw usage_error("Truncated gninatypes file " + path);
  return a;
}
inline void write_gninatypes(const std::string& path, const float* xyz, const int32_t* type, size_t n) {
  FILE* f = std::fopen(path.c_str(), "wb");
  if (!f) throw usage_error("Could not open " + path + " for writing");
  for (size_t i = 0; i < n; i++) {
    const float c[3] = {xyz[3 * i], xyz[3 * i + 1], xyz[3 * i + 2]};
    if (std::fwrite(c, 4, 3, f) != 3 || std::fwrite(&type[i], 4, 1, f) != 1) { std::fclose(f); throw internal_error("short write " + path); }
  }
  std::fclose(f);
}

// ---- pose queue in front of the batch entry point -------------------------------------------------------------
// add() copies one pose and returns its ticket (0, 1, 2, ... in submission order); when `capacity` poses are queued,
// or on flush() / destruction, ONE batch call scores them and `deliver` is invoked once per pose in ticket order.
// All poses of a batcher use either their own ligand centre (cnn_center unset, the reference default) or one fixed
// centre (cnnopts.cnn_center, cnn_torch_scorer.cpp:118-121).
class PoseBatcher {
 public:
  using Runner = std::function<Scores(const float* xyz, const int32_t* type, const int32_t* offsets, int n_poses,
                                      const float* centers)>;
  using Deliver = std::function<void(size_t ticket, float score, float affinity, float loss, float variance)>;

  PoseBatcher(Runner run, size_t capacity, Deliver deliver, const float* fixed_center = nullptr)
      : run_(std::move(run)), deliver_(std::move(deliver)), capacity_(capacity ? capacity : 1) {
    if (fixed_center) { has_center_ = true; center_ = {fixed_center[0], fixed_center[1], fixed_center[2]}; }
    offs_.push_back(0);
  }
  PoseBatcher(CNNScorer& s, size_t capacity, Deliver deliver, const float* fixed_center = nullptr)
      : PoseBatcher([&s](const float* x, const int32_t* t, const int32_t* o, int n, const float* c) {
                      return s.score_batch(x, t, o, n, c);
                    }, capacity, std::move(deliver), fixed_center) {}
  PoseBatcher(const PoseBatcher&) = delete;
  PoseBatcher& operator=(const PoseBatcher&) = delete;
  ~PoseBatcher() {
    try { flush(); } catch (...) {}  // destructors must not throw; call flush() yourself to see errors
  }

  size_t add(const float* xyz, const int32_t* type, int n_atoms) {
    xyz_.insert(xyz_.end(), xyz, xyz + 3 * (size_t)n_atoms);
    type_.insert(type_.end(), type, type + n_atoms);
    offs_.push_back((int32_t)type_.size());
    const size_t ticket = next_ticket_++;
    if (queued() >= capacity_) flush();
    return ticket;
  }
  size_t queued() const { return offs_.size() - 1; }
  size_t batches_run() const { return batches_; }
  void flush() {
    const size_t n = queued();
    if (n == 0) return;
    std::vector<float> centers;
    if (has_center_)
      for (size_t i = 0; i < n; i++) centers.insert(centers.end(), center_.begin(), center_.end());
    // take the queue first: a throwing runner must not leave half-delivered poses queued for a second delivery
    std::vector<float> xyz; std::vector<int32_t> type, offs;
    xyz.swap(xyz_); type.swap(type_); offs.swap(offs_);
    offs_.push_back(0);
    const size_t first = next_ticket_ - n;
    const Scores r = run_(xyz.data(), type.data(), offs.data(), (int)n, has_center_ ? centers.data() : nullptr);
    batches_++;
    for (size_t i = 0; i < n; i++) deliver_(first + i, r.score[i], r.affinity[i], r.loss[i], r.variance[i]);
  }

 private:
  Runner run_;
  Deliver deliver_;
  size_t capacity_, next_ticket_ = 0, batches_ = 0;
  bool has_center_ = false;
  std::array<float, 3> center_{};
  std::vector<float> xyz_;
  std::vector<int32_t> type_, offs_;
};

// grid_dim (lib/grid_dim.h:30-43) for the two out-of-box penalties
struct GridDim { float begin = 0, end = 0; int n = 0; };
using GridDims = std::array<GridDim, 3>;

class VinaScorer;
// per-atom empirical term used by NonCacheCNN's mixing (defined after VinaScorer)
// Scorer: score(xyz, types, n, want_gradient, affinity, loss, variance, gradient*) and info(model) like gb::CNNScorer; Emp:
// noncache_atoms(xyz, types, n, begin, end, cap, e, deriv) like gb::VinaScorer.  The class is a template so that its logic -- the
// part of non_cache_cnn.cpp that is host code here -- can be run against the REFERENCE's non_cache_cnn with test doubles on a
// machine without a GPU (oracle/ref_driver.cpp gref_noncache_cnn_*, tests/test_oracle_vs_reference_build.py); gb::NonCacheCNN below
// is the product instantiation.
template <class Scorer, class Emp>
class NonCacheCNNT {
  Scorer& scorer_;
  GridDims gd_, cnn_gd_;
  float slope_;
  // cnn_options::mix_emp_force / mix_emp_energy / empirical_weight (lib/non_cache_cnn.cpp:113-166)
  Emp* vina_ = nullptr;
  float emp_weight_ = 1.f;
  bool mix_force_ = false, mix_energy_ = false;
  bool reference_force_routing_ = true;

  static bool is_hydrogen(int32_t t) { return t == 0 || t == 1; }
  // non_cache::check_bounds_deriv, lib/non_cache.cpp:102-123
  float check_bounds(const GridDims& dims, const float* a, float* deriv) const {
    float pen = 0;
    for (int j = 0; j < 3; j++) {
      if (dims[j].n > 0) {
        if (a[j] < dims[j].begin) { if (deriv) deriv[j] += -1 * slope_; pen += std::fabs(a[j] - dims[j].begin); }
        else if (a[j] > dims[j].end) { if (deriv) deriv[j] += 1 * slope_; pen += std::fabs(a[j] - dims[j].end); }
      }
    }
    return pen * slope_;
  }

 public:
  // search box gd; the CNN box is centred on `cnn_center` with the model's dimension (set_bounding_box,
  // cnn_torch_scorer.cpp:229-241)
  NonCacheCNNT(Scorer& s, const GridDims& gd, const float cnn_center[3], float slope) : scorer_(s), gd_(gd), slope_(slope) {
    const gb_model_info inf = s.info(0);
    for (int i = 0; i < 3; i++) {
      cnn_gd_[i].begin = cnn_center[i] - inf.dimension / 2.0f;
      cnn_gd_[i].end = cnn_center[i] + inf.dimension / 2.0f;
      cnn_gd_[i].n = (int)(inf.dimension / inf.resolution);
    }
  }
  // --cnn_mix_emp_force / --cnn_mix_emp_energy / --cnn_empirical_weight: the empirical (smina) term evaluated directly over
  // the receptor (vina must have the receptor set) is blended into the forces / the energy of eval_deriv
  void set_empirical(Emp* vina, float weight, bool mix_force, bool mix_energy) {
    vina_ = vina; emp_weight_ = weight; mix_force_ = mix_force && vina; mix_energy_ = mix_energy && vina;
  }
  // true (default): the CNN forces are routed as CNNTorchScorer::score leaves them in the model -- getGradient's by-atom list consumed
  // compactly over the non-hydrogen atoms by model::add_minus_forces (cnn_torch_scorer.cpp:209-227, model.cu:247-259): the j-th heavy
  // atom receives entry j.  Identity without hydrogens.  false: the true per-atom gradient.
  void set_reference_force_routing(bool on) { reference_force_routing_ = on; }
  // eval (minus_forces == nullptr, lib/non_cache_cnn.cpp:33-54): loss + penalties.
  // eval_deriv (:79-169): additionally fills minus_forces[n_atoms][3] (zero for hydrogens); v = curl cap of the empirical
  // term.  With mix_emp_force the forces are (cnn + oob + w (emp + oob_search_box)) / (1 + w); with mix_emp_energy the
  // energy is (loss + penalties + w emp) / (1 + w) -- the identity test/gnina/test_min.py:45-61 checks to 1e-3.
  float eval(const float* lig_xyz, const int32_t* lig_type, int n, std::vector<float>* minus_forces = nullptr, float v = 1000.f) {
    float e = 0, aff = 0, loss = 0, var = 0;
    std::vector<float> grad;
    scorer_.score(lig_xyz, lig_type, n, minus_forces != nullptr, aff, loss, var, minus_forces ? &grad : nullptr);
    // eval_deriv starts from the loss and adds the penalties atom by atom (:100-111); eval sums the penalties first and adds the loss
    // last (:36-52) -- the same float association here
    if (minus_forces) e += loss;
    std::vector<float> emp_e, emp_d;
    const bool mixing = minus_forces && mix_force_;
    if (mixing) {
      const float b[3] = {gd_[0].begin, gd_[1].begin, gd_[2].begin}, en[3] = {gd_[0].end, gd_[1].end, gd_[2].end};
      vina_->noncache_atoms(lig_xyz, lig_type, n, b, en, v, emp_e, emp_d);
    }
    if (minus_forces) minus_forces->assign(3 * (size_t)n, 0.f);
    int n_heavy_before = 0;
    for (int i = 0; i < n; i++) {
      if (is_hydrogen(lig_type[i])) continue;
      const int src = reference_force_routing_ ? n_heavy_before : i;   // which entry of the by-atom gradient atom i receives
      n_heavy_before++;
      if (lig_type[i] < 0 || lig_type[i] >= 28) continue;
      float d_emp_box[3] = {0, 0, 0}, d_cnn_box[3] = {0, 0, 0};
      float pen = check_bounds(gd_, lig_xyz + 3 * i, minus_forces ? d_emp_box : nullptr);
      pen += check_bounds(cnn_gd_, lig_xyz + 3 * i, minus_forces ? d_cnn_box : nullptr);
      e += pen;
      if (minus_forces) {
        for (int k = 0; k < 3; k++) {
          float f = grad[3 * src + k] + (d_emp_box[k] + d_cnn_box[k]);   // the reference adds (0 + oob + cnn_oob) to the CNN force
          if (mixing) f = (f + emp_weight_ * (emp_d[3 * i + k] + d_emp_box[k])) / (1.0f + emp_weight_);
          (*minus_forces)[3 * i + k] = f;
        }
        if (mixing && mix_energy_) e += emp_weight_ * emp_e[i];
      }
    }
    if (!minus_forces) e += loss;
    if (minus_forces && mix_energy_) e /= (1.0f + emp_weight_);
    return e;
  }
};

class VinaScorer {
  gb_vina* h_ = nullptr;

 public:
  explicit VinaScorer(int device = 0, const float* weights6 = nullptr, float factor = 32.f) {
    check(gb_vina_create(device, weights6, factor, &h_));
  }
  VinaScorer(const VinaScorer&) = delete;
  ~VinaScorer() { gb_vina_destroy(h_); }
  void set_receptor(const float* xyz, const int32_t* t, int n) { check(gb_vina_set_receptor(h_, xyz, t, n)); }
  void cache_build(const float begin[3], const float end[3], const int32_t n[3], const std::vector<int32_t>& types) {
    check(gb_vina_cache_build(h_, begin, end, n, types.data(), (int)types.size()));
  }
  std::vector<float> cache_eval(const float* xyz, const int32_t* t, const int32_t* offs, int n_poses, float slope, float v,
                                std::vector<float>* deriv = nullptr) {
    std::vector<float> e(n_poses);
    if (deriv) deriv->assign(3 * (size_t)offs[n_poses], 0.f);
    check(gb_vina_cache_eval(h_, xyz, t, offs, n_poses, slope, v, e.data(), deriv ? deriv->data() : nullptr));
    return e;
  }
  std::vector<float> affinity(const float* xyz, const int32_t* t, const int32_t* offs, int n_poses, const float* num_tors,
                              float v = 1000.f) {
    std::vector<float> a(n_poses);
    check(gb_vina_score_exact(h_, xyz, t, offs, n_poses, num_tors, v, nullptr, a.data()));
    return a;
  }

  // the docking branch's final score (main/main.cpp:340-344): non_cache::eval on the search box with the search's tables + num_tors_div
  std::vector<float> docking_affinity(const float* xyz, const int32_t* t, const int32_t* offs, int n_poses, const float* num_tors,
                                      const float box_begin[3], const float box_end[3], float v = 1000.f, float slope = 1e3f) {
    std::vector<float> a(n_poses);
    check(gb_vina_score_noncache(h_, xyz, t, offs, n_poses, num_tors, v, slope, box_begin, box_end, nullptr, a.data()));
    return a;
  }

  // ---- docking inner loop (model::set / eval_deriv, quasi_newton, parallel_mc) --------------------------------------
  void set_ligand(const gb_ligand_topology& t) {
    check(gb_vina_set_ligand(h_, &t));
    n_tors_ = t.n_segments - 1; n_atoms_ = t.n_atoms;
  }
  int conf_size() const { return 7 + n_tors_; }  // position 3, orientation quaternion 4, torsions
  // energy (+ change[6+T] per conformation, + coordinates) of n conformations
  std::vector<float> eval_deriv(const float* confs, int n, const float v3[3], float slope, std::vector<float>* change = nullptr,
                                std::vector<float>* coords = nullptr) {
    std::vector<float> e(n);
    if (change) change->assign((size_t)n * (6 + n_tors_), 0.f);
    if (coords) coords->assign((size_t)n * 3 * n_atoms_, 0.f);
    check(gb_vina_eval_deriv(h_, confs, n, v3, slope, e.data(), change ? change->data() : nullptr, coords ? coords->data() : nullptr));
    return e;
  }
  std::vector<float> bfgs(float* confs_inout, int n, int maxiters, const float v3[3], float slope) {
    std::vector<float> e(n);
    check(gb_vina_bfgs(h_, confs_inout, n, maxiters, v3, slope, e.data(), nullptr, nullptr));
    return e;
  }
  // quasi_newton with the --minimize parameters (BFGSAccurateLineSearch, --minimize_early_term; lib/common.h:50-60)
  std::vector<float> minimize(float* confs_inout, int n, const gb_minimization_params& mp, const float v3[3], float slope) {
    std::vector<float> e(n);
    check(gb_vina_minimize(h_, confs_inout, n, &mp, v3, slope, e.data(), nullptr, nullptr));
    return e;
  }
  // model::eval_deriv with ig = non_cache (lib/non_cache.cpp:126-174): direct receptor sums, box = [begin, end]
  std::vector<float> eval_deriv_noncache(const float* confs, int n, const float v3[3], float slope, const float begin[3],
                                         const float end[3], std::vector<float>* change = nullptr) {
    std::vector<float> e(n), g((size_t)n * (6 + n_tors_));
    check(gb_vina_eval_deriv_noncache(h_, confs, n, v3, slope, begin, end, e.data(), g.data()));
    if (change) *change = std::move(g);
    return e;
  }
  // refine_structure (main/main.cpp:131-171) on n poses at once; e = max float for a pose that never entered the box
  std::vector<float> refine(float* confs_inout, int n, int maxiters, const float v3[3], const float begin[3], const float end[3],
                            std::vector<int32_t>* within = nullptr) {
    std::vector<float> e(n);
    std::vector<int32_t> ok(n);
    check(gb_vina_refine(h_, confs_inout, n, maxiters, v3, begin, end, e.data(), ok.data(), nullptr));
    if (within) *within = std::move(ok);
    return e;
  }
  // the --minimize / --local_only branch's refine_structure (main/main.cpp:264-268): the user's minimization_params
  std::vector<float> refine_minimize(float* confs_inout, int n, const gb_minimization_params& mp, const float v3[3], const float begin[3],
                                     const float end[3], std::vector<int32_t>* within = nullptr) {
    std::vector<float> e(n);
    std::vector<int32_t> ok(n);
    check(gb_vina_refine_minimize(h_, confs_inout, n, &mp, v3, begin, end, e.data(), ok.data(), nullptr));
    if (within) *within = std::move(ok);
    return e;
  }
  // per-atom empirical term of non_cache_cnn::eval_deriv (lib/non_cache_cnn.cpp:113-140)
  void noncache_atoms(const float* xyz, const int32_t* type, int n, const float begin[3], const float end[3], float cap,
                      std::vector<float>& e, std::vector<float>& deriv) {
    e.assign(n, 0.f); deriv.assign(3 * (size_t)n, 0.f);
    check(gb_vina_noncache_atoms(h_, xyz, type, n, begin, end, cap, e.data(), deriv.data()));
  }
  struct ChainOutputs {       // the chains' output_containers, flattened
    int n_chains = 0, S = 0;  // S = num_saved_mins
    std::vector<float> e, conf;
    std::vector<int32_t> n_out;
  };
  // parallel_mc::operator(): every chain of a ligand in one launch
  ChainOutputs parallel_mc(const gb_mc_params& P, const float corner1[3], const float corner2[3], const std::vector<uint32_t>& seeds,
                           float slope = 1e3f) {
    ChainOutputs o;
    o.n_chains = (int)seeds.size(); o.S = P.num_saved_mins;
    o.e.assign((size_t)o.n_chains * o.S, 0.f);
    o.conf.assign((size_t)o.n_chains * o.S * conf_size(), 0.f);
    o.n_out.assign(o.n_chains, 0);
    check(gb_vina_mc(h_, &P, corner1, corner2, seeds.data(), o.n_chains, slope, o.e.data(), o.conf.data(), o.n_out.data()));
    return o;
  }
  // merge_output_containers (host only): flat indices chain * S + k of the kept poses, best first
  static std::vector<int32_t> merge_outputs(const float* e, const float* coords, const int32_t* n_out, int n_chains, int S, int n_atoms,
                                            int max_size, float min_rmsd = 2.f) {
    std::vector<int32_t> kept(max_size > 0 ? max_size : 1);
    int32_t n = 0;
    check(gb_vina_merge_outputs(e, coords, n_out, n_chains, S, n_atoms, min_rmsd, max_size, kept.data(), &n));
    kept.resize(n);
    return kept;
  }

 private:
  int n_tors_ = 0, n_atoms_ = 0;
};

// non_cache_cnn (lib/non_cache_cnn.h) over the library's CNN scorer and, for --cnn_mix_emp_*, its Vina scorer
using NonCacheCNN = NonCacheCNNT<CNNScorer, VinaScorer>;

}  // namespace gb
