// Stand-in for libmolgrid (third-party, fetched unpinned by the reference's CMakeLists.txt:144-153, absent from /root/reference),
// written for oracle/_ref only so that the reference's OWN lib/torch_model.cpp and lib/cnn_torch_scorer.cpp compile where they lie.
// Unlike the Boost stand-ins this one carries arithmetic -- but none of its own: every number comes from oracle/gridmaker_ref.c
// (liboracle.so: the restatement of libmolgrid's published algorithm, pinned by the reference's binmap goldens), called through its C
// symbols.  What the compiled reference adds on top is therefore exactly the reference's own code: make_coordset, the centre choice,
// the rec+lig merge, the head post-processing, the ensemble arithmetic and the gradient routing into the model.
#pragma once
#include <cuda_runtime.h>
#include <cstdint>
#include <cstring>
#include <memory>
#include <sstream>
#include <stdexcept>
#include <string>
#include <utility>
#include <vector>

extern "C" {
int gbo_parse_typemap(const char* text, int* type_to_channel);
float gbo_smina_radius(int t);
const char* gbo_smina_name(int t);
void gbo_center(const float* xyz, int n, float* center);
int gbo_grid_npts(float resolution, float dimension);
void gbo_grid_forward(const float center[3], float resolution, float dimension, float radius_scale, int n_atoms, const float* xyz,
                      const int32_t* channel, const float* radius, int n_channels, float* out);
void gbo_grid_backward(const float center[3], float resolution, float dimension, float radius_scale, int n_atoms, const float* xyz,
                       const int32_t* channel, const float* radius, int n_channels, const float* gridgrad, float* atom_grad);
}

namespace libmolgrid {

struct random_engine_t { unsigned s = 0; void seed(unsigned v) { s = v; } };
inline random_engine_t random_engine;

// dense views over caller-owned memory (the reference only constructs them and passes them on)
template <class T, int N, bool isCUDA> struct Grid;
template <class T, bool isCUDA> struct Grid<T, 1, isCUDA> {
  T* p; size_t n0;
  T& operator[](size_t i) const { return p[i]; }
};
template <class T, bool isCUDA> struct Grid<T, 2, isCUDA> {
  T* p; size_t n0, n1;
  Grid(T* q, size_t a, size_t b) : p(q), n0(a), n1(b) {}
  Grid<T, 1, isCUDA> operator[](size_t i) const { return Grid<T, 1, isCUDA>{p + i * n1, n1}; }
};
template <class T, bool isCUDA> struct Grid<T, 4, isCUDA> {
  T* p; size_t n0, n1, n2, n3;
  Grid(T* q, size_t a, size_t b, size_t c, size_t d) : p(q), n0(a), n1(b), n2(c), n3(d) {}
};
struct MGrid2f {
  std::shared_ptr<std::vector<float>> buf; size_t n0, n1;
  MGrid2f(size_t a, size_t b) : buf(new std::vector<float>(a * b, 0.f)), n0(a), n1(b) {}
  Grid<float, 2, false> cpu() { return Grid<float, 2, false>(buf->data(), n0, n1); }
  Grid<float, 2, true> gpu() { return Grid<float, 2, true>(buf->data(), n0, n1); }   // oracle/_ref runs on the host only
};

struct AtomTyper {
  virtual ~AtomTyper() {}
  virtual unsigned num_types() const = 0;
  virtual std::pair<int, float> get_int_type(int smina_type) const = 0;
};
struct GninaIndexTyper {
  static std::string gnina_type_name(int t) { const char* s = gbo_smina_name(t); return s ? s : "Unknown"; }
};
// one channel per non-empty line of the map; the radius is the ORIGINAL smina type's xs radius
struct FileMappedGninaTyper : AtomTyper {
  int t2c[64]; int n = 0;
  explicit FileMappedGninaTyper(std::istream& in) {
    std::stringstream ss; ss << in.rdbuf();
    n = gbo_parse_typemap(ss.str().c_str(), t2c);
    if (n < 0) throw std::invalid_argument("unknown atom type name in map");
  }
  unsigned num_types() const override { return (unsigned)n; }
  std::pair<int, float> get_int_type(int t) const override { return std::make_pair((t >= 0 && t < 28) ? t2c[t] : -1, gbo_smina_radius(t)); }
};

struct CoordinateSet {
  std::vector<float> xyz; std::vector<int32_t> channel; std::vector<float> radius; unsigned max_type = 0;
  CoordinateSet(const std::vector<float3>& c, const std::vector<float>& types, const std::vector<float>& radii, unsigned ntypes)
      : radius(radii), max_type(ntypes) {
    for (const float3& p : c) { xyz.push_back(p.x); xyz.push_back(p.y); xyz.push_back(p.z); }
    for (float t : types) channel.push_back((int32_t)t);
  }
  // receptor channels first, ligand channels after them; untyped atoms (-1) stay untyped
  CoordinateSet(const CoordinateSet& rec, const CoordinateSet& lig) : xyz(rec.xyz), channel(rec.channel), radius(rec.radius) {
    xyz.insert(xyz.end(), lig.xyz.begin(), lig.xyz.end());
    radius.insert(radius.end(), lig.radius.begin(), lig.radius.end());
    for (int32_t c : lig.channel) channel.push_back(c < 0 ? -1 : c + (int32_t)rec.max_type);
    max_type = rec.max_type + lig.max_type;
  }
  size_t size() const { return channel.size(); }
  unsigned num_types() const { return max_type; }
  float3 center() const { float c[3]; gbo_center(xyz.data(), (int)size(), c); return make_float3(c[0], c[1], c[2]); }
};

struct Transform {
  Transform(float3, float, bool rotate) { if (rotate) throw std::runtime_error("oracle/_ref: libmolgrid's random rotations are not restated"); }
  void forward(const CoordinateSet&, CoordinateSet&) {}
  template <class G> void backward(const G&, G&, bool) {}
};

class GridMaker {
  float res_ = 0.5f, dim_ = 23.5f, rscale_ = 1.f;
 public:
  void initialize(float resolution, float dimension, bool binary, float rscale) {
    if (binary) throw std::runtime_error("oracle/_ref: binary grids are not restated");
    res_ = resolution; dim_ = dimension; rscale_ = rscale;
  }
  float get_dimension() const { return dim_; }
  float get_resolution() const { return res_; }
  long get_first_dim() const { return gbo_grid_npts(res_, dim_); }
  template <bool C> void forward(float3 c, const CoordinateSet& s, Grid<float, 4, C>& out) const {
    const float cc[3] = {c.x, c.y, c.z};
    gbo_grid_forward(cc, res_, dim_, rscale_, (int)s.size(), s.xyz.data(), s.channel.data(), s.radius.data(), (int)out.n0, out.p);
  }
  template <bool C> void backward(float3 c, const CoordinateSet& s, const Grid<float, 4, C>& g, Grid<float, 2, C>& atom_grad) const {
    const float cc[3] = {c.x, c.y, c.z};
    std::memset(atom_grad.p, 0, sizeof(float) * atom_grad.n0 * atom_grad.n1);
    gbo_grid_backward(cc, res_, dim_, rscale_, (int)s.size(), s.xyz.data(), s.channel.data(), s.radius.data(), (int)g.n0, g.p, atom_grad.p);
  }
};

}  // namespace libmolgrid
