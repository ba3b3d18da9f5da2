// Model blob parsing, atom typing tables (G0) and fp32-path weight upload.
// Reference: TorchModel ctor, gninasrc/lib/torch_model.cpp:49-118; make_coordset :120-142.
#include <cmath>
#include <cstring>
#include "gb_internal.h"

namespace gb {

// smina type names / xs_radius: gninasrc/lib/atom_constants.h:45-75 (enum), :101-133 (default_data)
const char* const kSminaNames[kNumSminaTypes] = {
    "Hydrogen", "PolarHydrogen", "AliphaticCarbonXSHydrophobe", "AliphaticCarbonXSNonHydrophobe",
    "AromaticCarbonXSHydrophobe", "AromaticCarbonXSNonHydrophobe", "Nitrogen", "NitrogenXSDonor",
    "NitrogenXSDonorAcceptor", "NitrogenXSAcceptor", "Oxygen", "OxygenXSDonor", "OxygenXSDonorAcceptor",
    "OxygenXSAcceptor", "Sulfur", "SulfurAcceptor", "Phosphorus", "Fluorine", "Chlorine", "Bromine",
    "Iodine", "Magnesium", "Manganese", "Zinc", "Calcium", "Iron", "GenericMetal", "Boron"};
const float kSminaXsRadius[kNumSminaTypes] = {0.37f, 0.37f, 1.9f, 1.9f, 1.9f, 1.9f, 1.8f, 1.8f, 1.8f, 1.8f,
                                              1.7f,  1.7f,  1.7f, 1.7f, 2.0f, 2.0f, 2.1f, 1.5f, 1.8f, 2.0f,
                                              2.2f,  1.2f,  1.2f, 1.2f, 1.2f, 1.2f, 1.2f, 1.92f};

void TypeMap::parse(const std::string& text) {
  for (int t = 0; t < kNumSminaTypes; t++) t2c[t] = -1;
  n_channels = 0;
  size_t pos = 0;
  while (pos < text.size()) {
    size_t eol = text.find('\n', pos);
    if (eol == std::string::npos) eol = text.size();
    bool used = false;
    size_t i = pos;
    while (i < eol) {
      while (i < eol && isspace((unsigned char)text[i])) i++;
      size_t s = i;
      while (i < eol && !isspace((unsigned char)text[i])) i++;
      if (i > s) {
        std::string name = text.substr(s, i - s);
        int found = -1;
        for (int t = 0; t < kNumSminaTypes; t++)
          if (name == kSminaNames[t]) found = t;
        if (found < 0) throw Error(GB_ERR_USAGE, "unknown atom type name in type map: " + name);
        t2c[found] = n_channels;
        used = true;
      }
    }
    if (used) n_channels++;
    pos = eol + 1;
  }
}

const HostTensor& Model::t(const std::string& n) const {
  auto it = tensors.find(n);
  if (it == tensors.end()) throw Error(GB_ERR_USAGE, "model " + name + " lacks tensor " + n);
  return it->second;
}

Model::~Model() {
  for (float* p : dev_allocs) cudaFree(p);
}

static float* upload(Model& m, const std::vector<float>& h) {
  float* d = nullptr;
  GB_CUDA(cudaMalloc(&d, h.size() * sizeof(float)));
  m.dev_allocs.push_back(d);
  GB_CUDA(cudaMemcpy(d, h.data(), h.size() * sizeof(float), cudaMemcpyHostToDevice));
  return d;
}

static void add_conv(Model& m, const std::string& key, const std::string& bn_key = "") {
  const HostTensor& w = m.t(key + ".weight");
  const HostTensor& b = m.t(key + ".bias");
  GB_CHECK(w.shape.size() == 5 && w.shape[2] == w.shape[3] && w.shape[3] == w.shape[4], "conv weight shape");
  ConvF32 c;
  c.cout = w.shape[0]; c.cin = w.shape[1]; c.ks = w.shape[2];
  int k3 = c.ks * c.ks * c.ks;
  std::vector<float> r((size_t)c.cin * k3 * c.cout);
  for (int co = 0; co < c.cout; co++)
    for (int ci = 0; ci < c.cin; ci++)
      for (int k = 0; k < k3; k++) r[((size_t)ci * k3 + k) * c.cout + co] = w.data[((size_t)co * c.cin + ci) * k3 + k];
  c.w = upload(m, r);
  c.bias = upload(m, std::vector<float>(b.data, b.data + b.nelem));
  // backward-data: dX[ci] = sum_{co,tap} dY[co](pos - tap) W[co][ci][tap]  ==  conv(dY, W'), W'[ci][co][tap] = W[co][ci][flip(tap)]
  std::vector<float> rt((size_t)c.cout * k3 * c.cin);
  for (int co = 0; co < c.cout; co++)
    for (int ci = 0; ci < c.cin; ci++)
      for (int k = 0; k < k3; k++)
        rt[((size_t)co * k3 + (k3 - 1 - k)) * c.cin + ci] = w.data[((size_t)co * c.cin + ci) * k3 + k];
  c.wT = upload(m, rt);
  c.zero_bias = upload(m, std::vector<float>(c.cin, 0.f));
  if (!bn_key.empty()) {
    // eval-mode BatchNorm3d, eps 1e-5 (read from the dense graphs): y = (x-mean)/sqrt(var+eps)*w + b
    const HostTensor &g = m.t(bn_key + ".weight"), &be = m.t(bn_key + ".bias"), &mu = m.t(bn_key + ".running_mean"),
                     &var = m.t(bn_key + ".running_var");
    GB_CHECK((int)g.nelem == c.cin, "bn size");
    std::vector<float> sc(c.cin), sh(c.cin);
    for (int i = 0; i < c.cin; i++) {
      float inv = 1.0f / std::sqrt(var.data[i] + 1e-5f);
      sc[i] = g.data[i] * inv;
      sh[i] = be.data[i] - mu.data[i] * sc[i];
    }
    c.bn_scale = upload(m, sc);
    c.bn_shift = upload(m, sh);
  }
  m.convs[key] = c;
}

Model* load_model_from_memory(const void* data, size_t n, int device, const std::string& label) {
  const std::string bad = "Could not read torch model " + label;  // torch_model.cpp:116
  if (n < 48 || memcmp(data, "GNB200W1", 8) != 0) throw Error(GB_ERR_USAGE, bad);
  std::unique_ptr<Model> m(new Model);
  m->device = device;
  m->raw.assign((const char*)data, (const char*)data + n);
  const char* p = m->raw.data();
  uint32_t arch, nt, flags, nl, rl, ll;
  memcpy(&arch, p + 8, 4); memcpy(&nt, p + 12, 4);
  memcpy(&m->resolution, p + 16, 4); memcpy(&m->dimension, p + 20, 4); memcpy(&m->radius_scaling, p + 24, 4);
  memcpy(&flags, p + 28, 4); memcpy(&nl, p + 32, 4); memcpy(&rl, p + 36, 4); memcpy(&ll, p + 40, 4);
  size_t off = 48;
  if (off + nl + rl + ll > n) throw Error(GB_ERR_USAGE, bad);
  m->name.assign(p + off, nl); off += nl;
  m->recmap.assign(p + off, rl); off += rl;
  m->ligmap.assign(p + off, ll); off += ll;
  off += (8 - off % 8) % 8;
  m->arch = (int)arch;
  m->apply_logistic_loss = flags & 1; m->skip_softmax = flags & 2;
  if (arch < 1 || arch > 4) throw Error(GB_ERR_USAGE, bad);
  const size_t entry = 96 + 4 + 24 + 4 + 8 + 8;
  if (off + entry * nt > n) throw Error(GB_ERR_USAGE, bad);
  for (uint32_t i = 0; i < nt; i++) {
    const char* e = p + off + entry * i;
    std::string tn(e, strnlen(e, 96));
    uint32_t ndim, dims[6]; uint64_t toff, nelem;
    memcpy(&ndim, e + 96, 4); memcpy(dims, e + 100, 24); memcpy(&toff, e + 128, 8); memcpy(&nelem, e + 136, 8);
    // a user-supplied blob (--cnn_models) is untrusted input: the tensor must lie inside the file (overflow-safe), be
    // float-aligned, and its dimensions must multiply to its element count
    if (ndim > 6 || toff > n || toff % 4 != 0 || nelem > (n - toff) / 4) throw Error(GB_ERR_USAGE, bad);
    uint64_t prod = 1;
    for (uint32_t d = 0; d < ndim; d++) {
      if (dims[d] == 0 || dims[d] > (1u << 28) || prod > (1ull << 40)) throw Error(GB_ERR_USAGE, bad);
      prod *= dims[d];
    }
    if (prod != nelem) throw Error(GB_ERR_USAGE, bad);
    HostTensor t;
    for (uint32_t d = 0; d < ndim; d++) t.shape.push_back((int)dims[d]);
    t.data = (const float*)(p + toff); t.nelem = nelem;
    m->tensors[tn] = t;
  }
  m->rec.parse(m->recmap);
  m->lig.parse(m->ligmap);
  m->n_channels = m->rec.n_channels + m->lig.n_channels;
  m->npts = (int)std::lround(m->dimension / m->resolution) + 1;

  GB_CUDA(cudaSetDevice(device));
  if (m->arch == GB_ARCH_OVERLAP) {  // no parameters: channel 0 = receptor, channel 1 = ligand
    if (m->n_channels != 2) throw Error(GB_ERR_USAGE, bad);
    return m.release();
  }
  if (m->arch == GB_ARCH_DEFAULT2018) {
    for (const char* k : {"unit1_conv", "unit2_conv", "unit3_conv", "unit4_conv", "unit5_conv"}) add_conv(*m, k);
  } else if (m->arch == GB_ARCH_DEFAULT2017) {
    for (const char* k : {"unit1_conv1", "unit2_conv1", "unit3_conv1"}) add_conv(*m, k);
  } else {
    add_conv(*m, "data_enc_init_conv");
    for (int L = 0; L < 3; L++) {
      for (int i = 0; i < 4; i++) {
        std::string base = "dense_block_" + std::to_string(L) + ".data_enc_level" + std::to_string(L);
        add_conv(*m, base + "_conv" + std::to_string(i), base + "_batchnorm_conv" + std::to_string(i));
      }
      if (L < 2) add_conv(*m, "data_enc_level" + std::to_string(L) + "_bottleneck");
    }
  }
  const HostTensor &pw = m->t("pose_output.weight"), &pb = m->t("pose_output.bias"), &aw = m->t("affinity_output.weight"),
                   &ab = m->t("affinity_output.bias");
  GB_CHECK(pw.shape.size() == 2 && pw.shape[0] == 2 && aw.shape[0] == 1 && aw.shape[1] == pw.shape[1], "head shapes");
  m->fc_features = pw.shape[1];
  std::vector<float> fw((size_t)3 * m->fc_features), fb(3);
  memcpy(fw.data(), pw.data, sizeof(float) * 2 * m->fc_features);
  memcpy(fw.data() + 2 * (size_t)m->fc_features, aw.data, sizeof(float) * m->fc_features);
  fb[0] = pb.data[0]; fb[1] = pb.data[1]; fb[2] = ab.data[0];
  m->fc_w = upload(*m, fw);
  m->fc_b = upload(*m, fb);
  return m.release();
}

}  // namespace gb
