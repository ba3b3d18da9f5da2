// gninasrc/lib/cnn_b200_scorer.h -- the ONE class a gnina maintainer adds to use libgnina_b200.so for CNN scoring.
// It implements gnina's own DLScorer interface (lib/dl_scorer.h:23-66) the way CNNTorchScorer does (lib/cnn_torch_scorer.cpp) and
// forwards the arithmetic to the C ABI of include/gnina_b200.h.  Compiled against the reference's real headers by
// tests/test_integration_adapters.py (with the Boost / OpenBabel stand-ins of oracle/ref_shim where those libraries are absent).
#ifndef CNN_B200_SCORER_H_
#define CNN_B200_SCORER_H_

#include <cmath>
#include <memory>
#include <stdexcept>
#include <string>
#include <vector>

#include <algorithm>
#include <filesystem>

#include "dl_scorer.h"
#include "gnina_b200.hpp"   // this repo's include/: the C ABI (gnina_b200.h) + gb::expand_model_names

class CNNB200Scorer : public DLScorer {
  gb_cnn* h = nullptr;
  std::vector<gb_model*> models;
  gb_model_info info0{};                       // grid dimension / resolution of the first model (get_grid_dim / get_grid_res); copies keep it
  int device = 0;
  std::string blob_dir;                        // where the converted model blobs live (tools/extract_models.py output)
  std::vector<float> uploaded_receptor;        // coordinates of the last gb_cnn_set_receptor: re-upload only when flex atoms moved
  std::vector<gfloat3> gradient;               // as CNNTorchScorer::gradient: indexed by movable-atom index
  std::vector<float> grad_lig, grad_rec;

  static void check(int rc) {
    if (rc == GB_ERR_USAGE) throw usage_error(gb_last_error());
    if (rc != GB_OK) throw std::runtime_error(gb_last_error());
  }

  // the built-in names (cnn_torch_scorer.cpp:28-62: the default ensemble, "fast", "default1.0", "<prefix>_ensemble") resolve to the
  // blobs converted from the TorchScript models the reference embeds: one <name>.gbw per model under blob_dir
  void load_models() {
    std::vector<std::string> builtin;
    for (const auto& f : std::filesystem::directory_iterator(blob_dir))
      if (f.path().extension() == ".gbw") builtin.push_back(f.path().stem().string());
    std::sort(builtin.begin(), builtin.end());
    std::vector<std::string> names = cnnopts.cnn_model_names;
    if (!names.empty() || cnnopts.cnn_models.empty())                            // no names and no files: the default ensemble
      for (const std::string& name : gb::expand_model_names(names, builtin)) {
        if (!std::binary_search(builtin.begin(), builtin.end(), name)) throw usage_error("Invalid model name: " + name);  // :71
        gb_model* m = nullptr;
        check(gb_model_load((blob_dir + "/" + name + ".gbw").c_str(), device, &m));
        models.push_back(m);
      }
    for (const std::string& f : cnnopts.cnn_models) {                            // external model files, :84-89
      gb_model* m = nullptr;
      check(gb_model_load(f.c_str(), device, &m));
      models.push_back(m);
    }
  }

  void upload_receptor_if_changed() {
    const size_t n = receptor_coords.size();
    const float* xyz = reinterpret_cast<const float*>(receptor_coords.data());
    if (uploaded_receptor.size() == 3 * n && std::equal(xyz, xyz + 3 * n, uploaded_receptor.begin())) return;
    check(gb_cnn_set_receptor(h, xyz, reinterpret_cast<const int32_t*>(receptor_smtypes.data()), (int)n));
    uploaded_receptor.assign(xyz, xyz + 3 * n);
  }

 public:
  CNNB200Scorer() {}
  CNNB200Scorer(const cnn_options& opts, int device_, const std::string& blob_dir_) : DLScorer(opts), device(device_), blob_dir(blob_dir_) {
    if (cnnopts.cnn_scoring == CNNnone) return;                                  // cnn_torch_scorer.cpp:25-26
    load_models();
    if (!models.empty()) check(gb_model_get_info(models[0], &info0));
    check(gb_cnn_create(models.data(), (int)models.size(), device, &h));
    check(gb_cnn_set_option(h, "cnn_rotation", (double)cnnopts.cnn_rotations));  // :127-163
    check(gb_cnn_set_option(h, "rotation_seed", (double)cnnopts.seed));
  }
  ~CNNB200Scorer() override {
    if (h) gb_cnn_destroy(h);
    for (gb_model* m : models) gb_model_release(m);
  }
  CNNB200Scorer(const CNNB200Scorer&) = delete;
  CNNB200Scorer& operator=(const CNNB200Scorer&) = delete;

  bool initialized() const override { return h != nullptr && gb_cnn_num_models(h) > 0; }
  bool has_affinity() const override { return true; }

  // CNNTorchScorer::score (lib/cnn_torch_scorer.cpp:105-198).  ALERT, as there: clears minus forces.
  float score(model& m, bool compute_gradient, float& affinity, float& loss, float& variance) override {
    boost::lock_guard<boost::recursive_mutex> guard(*mtx);
    if (!initialized()) return -1.0;
    setLigand(m);                                                                // DLScorer, unchanged (dl_scorer.cpp:36-88)
    setReceptor(m);                                                              // (:93-193)
    m.clear_minus_forces();
    upload_receptor_if_changed();
    const int32_t offs[2] = {0, (int32_t)ligand_coords.size()};
    const float center[3] = {(float)cnnopts.cnn_center[0], (float)cnnopts.cnn_center[1], (float)cnnopts.cnn_center[2]};
    const float* centers = std::isnan(cnnopts.cnn_center[0]) ? nullptr : center;  // NaN = recalculate from the ligand (:135-138)
    const float* lig_xyz = reinterpret_cast<const float*>(ligand_coords.data());
    const int32_t* lig_t = reinterpret_cast<const int32_t*>(ligand_smtypes.data());
    float s = 0;
    if (!compute_gradient) {
      check(gb_cnn_score_batch(h, lig_xyz, lig_t, offs, 1, centers, &s, &affinity, &loss, &variance));
      return s;
    }
    grad_lig.assign(3 * ligand_coords.size(), 0.f);
    grad_rec.assign(3 * receptor_coords.size(), 0.f);
    const bool flex = !receptor_map.empty();                                     // optimisation of flexible residues (:209-211)
    check(gb_cnn_score_grad(h, lig_xyz, lig_t, offs, 1, centers, &s, &affinity, &loss, &variance, grad_lig.data(),
                            flex ? grad_rec.data() : nullptr));
    // getGradient (:200-222): scatter into a vector indexed by movable-atom index; rotations and the 1 / cnt scaling of :176-179
    // have already been applied by the library
    gradient.assign(receptor_map.size() + ligand_map.size(), gfloat3(0, 0, 0));
    for (sz i = 0, n = ligand_map.size(); i < n; i++)
      gradient[ligand_map[i]] = gfloat3(grad_lig[3 * i], grad_lig[3 * i + 1], grad_lig[3 * i + 2]);
    for (sz i = 0, n = receptor_map.size(); i < n; i++)
      gradient[receptor_map[i]] = gfloat3(grad_rec[3 * i], grad_rec[3 * i + 1], grad_rec[3 * i + 2]);
    m.add_minus_forces(gradient);                                                // lib/model.cu:247-259 (hydrogens skipped)
    return s;
  }
  float score(model& m, float& variance) override {
    float aff = 0, loss = 0;
    return score(m, false, aff, loss, variance);
  }

  // fresh_copy (lib/cnn_torch_scorer.h:54): an independent handle for another thread that shares the device weights
  std::shared_ptr<DLScorer> fresh_copy() const override {
    std::shared_ptr<CNNB200Scorer> c = std::make_shared<CNNB200Scorer>();
    c->cnnopts = cnnopts; c->device = device; c->blob_dir = blob_dir; c->info0 = info0;
    if (h) check(gb_cnn_clone(h, &c->h));
    return c;
  }

  // lib/cnn_torch_scorer.cpp:229-241
  void set_bounding_box(grid_dims& box) const override {
    if (!h) return;
    const vec center = get_center();
    const fl dim = info0.dimension, n = dim / info0.resolution, half = dim / 2.0;
    for (unsigned i = 0; i < 3; i++) {
      box[i].begin = center[i] - half;
      box[i].end = center[i] + half;
      box[i].n = n;
    }
  }
};

#endif
