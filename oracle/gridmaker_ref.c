/*
 * oracle/gridmaker_ref.c — TEST INFRASTRUCTURE ONLY (CPU oracle, never on the product path).
 *
 * Scalar C restatement of the voxeliser the reference calls at gninasrc/lib/torch_model.cpp:108,181,203
 * (libmolgrid::GridMaker::initialize / forward / backward) and of the atom typing it feeds it
 * (make_coordset, gninasrc/lib/torch_model.cpp:120-142; FileMappedGninaTyper built from the model's
 * recmap/ligmap, torch_model.cpp:16-46,110-113).
 *
 * The arithmetic lives in a THIRD-PARTY dependency that is not under /root/reference:
 *   gnina/libmolgrid, fetched UNPINNED at configure time (CMakeLists.txt:144-153).
 * Its published algorithm (GridMaker::calc_point, Gaussian radius multiple G=1, final multiple 1.5):
 *   rho(d, r) = exp(-2 d^2 / r^2)                     d <= r
 *             = (A q + B) q + C,  q = d / r          r < d < 1.5 r   (A=4e^-2, B=-12e^-2, C=9e^-2)
 *             = 0                                     otherwise
 *   grid point (i,j,k) = center - dimension/2 + resolution*(i,j,k); layout out[c][i][j][k], k fastest.
 * Parity is PINNED for the forward density by the reference's own golden files
 *   test/gninagrid/files/cc_0.48.35.binmap, ccsmall_0.33.35.binmap (tolerance 1e-4,
 *   test/gninagrid/compare_bin.py:24) — see tests/test_oracle_gridmaker.py.
 * Backward (G2) and CoordinateSet::center() have no golden in the reference: "parity unpinned";
 * backward is checked against finite differences of the (pinned) forward.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define GB_NUM_SMINA_TYPES 28

/* smina type names and xs_radius: gninasrc/lib/atom_constants.h:45-75 (enum), :101-133 (default_data) */
static const char *const smina_names[GB_NUM_SMINA_TYPES] = {
    "Hydrogen", "PolarHydrogen", "AliphaticCarbonXSHydrophobe", "AliphaticCarbonXSNonHydrophobe",
    "AromaticCarbonXSHydrophobe", "AromaticCarbonXSNonHydrophobe", "Nitrogen", "NitrogenXSDonor",
    "NitrogenXSDonorAcceptor", "NitrogenXSAcceptor", "Oxygen", "OxygenXSDonor", "OxygenXSDonorAcceptor",
    "OxygenXSAcceptor", "Sulfur", "SulfurAcceptor", "Phosphorus", "Fluorine", "Chlorine", "Bromine",
    "Iodine", "Magnesium", "Manganese", "Zinc", "Calcium", "Iron", "GenericMetal", "Boron"};
static const float smina_xs_radius[GB_NUM_SMINA_TYPES] = {
    0.37f, 0.37f, 1.9f, 1.9f, 1.9f, 1.9f, 1.8f, 1.8f, 1.8f, 1.8f, 1.7f, 1.7f, 1.7f, 1.7f,
    2.0f,  2.0f,  2.1f, 1.5f, 1.8f, 2.0f, 2.2f, 1.2f, 1.2f, 1.2f, 1.2f, 1.2f, 1.2f, 1.92f};

const char *gbo_smina_name(int t) { return (t >= 0 && t < GB_NUM_SMINA_TYPES) ? smina_names[t] : 0; }
float gbo_smina_radius(int t) { return (t >= 0 && t < GB_NUM_SMINA_TYPES) ? smina_xs_radius[t] : 0.f; }

/* One channel per non-empty LINE of the map text; names on a line are merged into that channel
 * (torch_model.cpp:16-46 default maps; FileMappedGninaTyper semantics).  type_to_channel[t] = -1
 * for types that appear on no line.  Returns the number of channels, or -1 on an unknown name. */
int gbo_parse_typemap(const char *text, int *type_to_channel) {
  for (int t = 0; t < GB_NUM_SMINA_TYPES; t++) type_to_channel[t] = -1;
  int nch = 0;
  const char *p = text;
  while (*p) {
    const char *eol = strchr(p, '\n');
    size_t len = eol ? (size_t)(eol - p) : strlen(p);
    int used = 0;
    size_t i = 0;
    while (i < len) {
      while (i < len && (p[i] == ' ' || p[i] == '\t' || p[i] == '\r')) i++;
      size_t s = i;
      while (i < len && !(p[i] == ' ' || p[i] == '\t' || p[i] == '\r')) i++;
      if (i > s) {
        int found = -1;
        for (int t = 0; t < GB_NUM_SMINA_TYPES; t++)
          if (strlen(smina_names[t]) == i - s && strncmp(smina_names[t], p + s, i - s) == 0) found = t;
        if (found < 0) return -1;
        type_to_channel[found] = nch;
        used = 1;
      }
    }
    if (used) nch++;
    p += len + (eol ? 1 : 0);
  }
  return nch;
}

/* make_coordset (torch_model.cpp:120-142): channel = typer(smina type) or -1, radius = xs_radius of the
 * ORIGINAL smina type.  channel_offset implements the rec/lig merge (lig channels follow rec channels,
 * torch_model.cpp:168). */
void gbo_type_atoms(const int32_t *smina_type, int n, const int *type_to_channel, int channel_offset,
                    int32_t *channel, float *radius) {
  for (int i = 0; i < n; i++) {
    int t = smina_type[i];
    int c = (t >= 0 && t < GB_NUM_SMINA_TYPES) ? type_to_channel[t] : -1;
    channel[i] = c < 0 ? -1 : c + channel_offset;
    radius[i] = gbo_smina_radius(t);
  }
}

/* CoordinateSet::center(): arithmetic mean over ALL atoms passed (hydrogens included). Unpinned. */
void gbo_center(const float *xyz, int n, float *center) {
  float sx = 0, sy = 0, sz = 0;
  for (int i = 0; i < n; i++) { sx += xyz[3 * i]; sy += xyz[3 * i + 1]; sz += xyz[3 * i + 2]; }
  if (n > 0) { sx /= n; sy /= n; sz /= n; }
  center[0] = sx; center[1] = sy; center[2] = sz;
}

typedef struct { float A, B, C; } quad_t;
static quad_t quad_coeffs(void) {
  const float G = 1.0f; /* gaussian_radius_multiple */
  float e = expf(-2.f * G * G);
  quad_t q = {e * 4.f * G * G, -e * (4.f * G + 8.f * G * G * G), e * (4.f * G * G * G * G + 4.f * G * G + 1.f)};
  return q;
}

static inline float density(float dx, float dy, float dz, float ar, quad_t qc) {
  float rsq = dx * dx + dy * dy + dz * dz;
  float dist = sqrtf(rsq);
  if (dist >= ar * 1.5f) return 0.f;
  if (dist <= ar) return expf(-2.f * dist * dist / (ar * ar));
  float dr = dist / ar;
  float q = (qc.A * dr + qc.B) * dr + qc.C;
  return q > 0.f ? q : 0.f;
}

/* GridMaker::forward for one pose.  out must hold n_channels*N^3 floats and is OVERWRITTEN (the
 * reference zero-fills it first, torch_model.cpp:179).  N = round(dimension/resolution)+1. */
int gbo_grid_npts(float resolution, float dimension) { return (int)roundf(dimension / resolution) + 1; }

void gbo_grid_forward(const float center[3], float resolution, float dimension, float radius_scale, int n_atoms,
                      const float *xyz, const int32_t *channel, const float *radius, int n_channels, float *out) {
  const int N = gbo_grid_npts(resolution, dimension);
  const quad_t qc = quad_coeffs();
  memset(out, 0, sizeof(float) * (size_t)n_channels * N * N * N);
  const float ox = center[0] - dimension / 2.f, oy = center[1] - dimension / 2.f, oz = center[2] - dimension / 2.f;
  for (int a = 0; a < n_atoms; a++) {
    int c = channel[a];
    if (c < 0 || c >= n_channels) continue;
    float ar = radius[a] * radius_scale;
    float reach = ar * 1.5f;
    float ax = xyz[3 * a], ay = xyz[3 * a + 1], az = xyz[3 * a + 2];
    int lo[3], hi[3];
    const float o[3] = {ox, oy, oz}, p[3] = {ax, ay, az};
    int empty = 0;
    for (int d = 0; d < 3; d++) {
      lo[d] = (int)floorf((p[d] - reach - o[d]) / resolution);
      hi[d] = (int)ceilf((p[d] + reach - o[d]) / resolution);
      if (lo[d] < 0) lo[d] = 0;
      if (hi[d] > N - 1) hi[d] = N - 1;
      if (lo[d] > hi[d]) empty = 1;
    }
    if (empty) continue;
    float *g = out + (size_t)c * N * N * N;
    for (int i = lo[0]; i <= hi[0]; i++) {
      float dx = (ox + i * resolution) - ax;
      for (int j = lo[1]; j <= hi[1]; j++) {
        float dy = (oy + j * resolution) - ay;
        for (int k = lo[2]; k <= hi[2]; k++) {
          float dz = (oz + k * resolution) - az;
          g[((size_t)i * N + j) * N + k] += density(dx, dy, dz, ar, qc);
        }
      }
    }
  }
}

/* GridMaker::backward (torch_model.cpp:203): atom gradient of sum_v gridgrad[c_a][v]*rho(|v-x_a|, r_a).
 * d rho/d x_a = rho'(d) * (x_a - v)/d ; rho'(d) = -4 d/r^2 exp(-2d^2/r^2) (d<=r); (2 A q + B)/r (r<d<1.5r). */
void gbo_grid_backward(const float center[3], float resolution, float dimension, float radius_scale, int n_atoms,
                       const float *xyz, const int32_t *channel, const float *radius, int n_channels,
                       const float *gridgrad, float *atomgrad) {
  const int N = gbo_grid_npts(resolution, dimension);
  const quad_t qc = quad_coeffs();
  const float ox = center[0] - dimension / 2.f, oy = center[1] - dimension / 2.f, oz = center[2] - dimension / 2.f;
  for (int a = 0; a < n_atoms; a++) {
    atomgrad[3 * a] = atomgrad[3 * a + 1] = atomgrad[3 * a + 2] = 0.f;
    int c = channel[a];
    if (c < 0 || c >= n_channels) continue;
    float ar = radius[a] * radius_scale, reach = ar * 1.5f;
    float ax = xyz[3 * a], ay = xyz[3 * a + 1], az = xyz[3 * a + 2];
    const float o[3] = {ox, oy, oz}, p[3] = {ax, ay, az};
    int lo[3], hi[3], empty = 0;
    for (int d = 0; d < 3; d++) {
      lo[d] = (int)floorf((p[d] - reach - o[d]) / resolution);
      hi[d] = (int)ceilf((p[d] + reach - o[d]) / resolution);
      if (lo[d] < 0) lo[d] = 0;
      if (hi[d] > N - 1) hi[d] = N - 1;
      if (lo[d] > hi[d]) empty = 1;
    }
    if (empty) continue;
    const float *g = gridgrad + (size_t)c * N * N * N;
    double gx = 0, gy = 0, gz = 0;
    for (int i = lo[0]; i <= hi[0]; i++)
      for (int j = lo[1]; j <= hi[1]; j++)
        for (int k = lo[2]; k <= hi[2]; k++) {
          float dx = (ox + i * resolution) - ax, dy = (oy + j * resolution) - ay, dz = (oz + k * resolution) - az;
          float dist = sqrtf(dx * dx + dy * dy + dz * dz);
          if (dist >= reach || dist == 0.f) continue;
          float dr; /* d rho / d dist */
          if (dist <= ar) dr = -4.f * dist / (ar * ar) * expf(-2.f * dist * dist / (ar * ar));
          else dr = (2.f * qc.A * (dist / ar) + qc.B) / ar;
          float s = g[((size_t)i * N + j) * N + k] * dr / dist;
          /* d dist / d x_a = (x_a - v)/dist = -dx/dist */
          gx += -(double)s * dx; gy += -(double)s * dy; gz += -(double)s * dz;
        }
    atomgrad[3 * a] = (float)gx; atomgrad[3 * a + 1] = (float)gy; atomgrad[3 * a + 2] = (float)gz;
  }
}

/* Batched driver used only by bench.py's cpu_baseline leg: one receptor, many ligand poses, the receptor
 * re-voxelised for every pose exactly as torch_model.cpp:153-181 does.  pose_offsets has n_poses+1 entries.
 * centers: n_poses*3 or NULL (NULL => ligand mean, torch_model.cpp:163-166).  out: [n_poses][C][N^3].
 * n_threads pthreads, poses striped across them. */
#include <pthread.h>
typedef struct {
  float resolution, dimension, radius_scale;
  int n_rec; const float *rec_xyz; const int32_t *rec_channel; const float *rec_radius;
  const float *lig_xyz; const int32_t *lig_channel; const float *lig_radius; const int32_t *pose_offsets;
  int n_poses; const float *centers; int n_channels; float *out; int tid, n_threads;
} batch_job;

static void *batch_worker(void *arg) {
  batch_job *j = (batch_job *)arg;
  const int N = gbo_grid_npts(j->resolution, j->dimension);
  const size_t gsz = (size_t)j->n_channels * N * N * N;
  for (int p = j->tid; p < j->n_poses; p += j->n_threads) {
    int b = j->pose_offsets[p], e = j->pose_offsets[p + 1], nl = e - b;
    int n = j->n_rec + nl;
    float *xyz = (float *)malloc(sizeof(float) * 3 * n);
    int32_t *ch = (int32_t *)malloc(sizeof(int32_t) * n);
    float *rad = (float *)malloc(sizeof(float) * n);
    memcpy(xyz, j->rec_xyz, sizeof(float) * 3 * j->n_rec);
    memcpy(ch, j->rec_channel, sizeof(int32_t) * j->n_rec);
    memcpy(rad, j->rec_radius, sizeof(float) * j->n_rec);
    memcpy(xyz + 3 * j->n_rec, j->lig_xyz + 3 * b, sizeof(float) * 3 * nl);
    memcpy(ch + j->n_rec, j->lig_channel + b, sizeof(int32_t) * nl);
    memcpy(rad + j->n_rec, j->lig_radius + b, sizeof(float) * nl);
    float c[3];
    if (j->centers) { c[0] = j->centers[3 * p]; c[1] = j->centers[3 * p + 1]; c[2] = j->centers[3 * p + 2]; }
    else gbo_center(j->lig_xyz + 3 * b, nl, c);
    gbo_grid_forward(c, j->resolution, j->dimension, j->radius_scale, n, xyz, ch, rad, j->n_channels,
                     j->out + gsz * p);
    free(xyz); free(ch); free(rad);
  }
  return 0;
}

void gbo_grid_forward_batch(float resolution, float dimension, float radius_scale, int n_rec, const float *rec_xyz,
                            const int32_t *rec_channel, const float *rec_radius, const float *lig_xyz,
                            const int32_t *lig_channel, const float *lig_radius, const int32_t *pose_offsets,
                            int n_poses, const float *centers, int n_channels, float *out, int n_threads) {
  if (n_threads < 1) n_threads = 1;
  if (n_threads > 256) n_threads = 256;
  pthread_t th[256];
  batch_job jobs[256];
  for (int t = 0; t < n_threads; t++) {
    batch_job j = {resolution, dimension, radius_scale, n_rec, rec_xyz, rec_channel, rec_radius, lig_xyz,
                   lig_channel, lig_radius, pose_offsets, n_poses, centers, n_channels, out, t, n_threads};
    jobs[t] = j;
    pthread_create(&th[t], 0, batch_worker, &jobs[t]);
  }
  for (int t = 0; t < n_threads; t++) pthread_join(th[t], 0);
}
