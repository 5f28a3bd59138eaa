// Seam 2 of the drop-in boundary: the symbols the reference's UNMODIFIED CPython glue
// (src/python/c/*.c, src/support/c_ptrdict.c, c_logging.c) links against, supplied in C++ on top of
// the C ABI of libatomistica_b200.so.  In the reference these symbols come from Fortran
// (src/python/f90/{particles,neighbors}_wrap.f90, python_helper.f90, the generated
// potentials_factory_f90.f90 and src/support/{error,f_logging,atomistica}.f90); here
//   f_particles_*  / data accessors   -> host arrays + atx_particles_*
//   f_neighbors_*  / f_get_* / f_pack_* -> atx_neighbors_* (+ a host copy of the list on request)
//   potential_classes[]               -> python_<pot>_{new,free,register_data,init,bind_to,
//                                        energy_and_forces} over atx_bop_* / atx_eam_*
//   c_push_error_with_info, get_full_error_string, c_prlog, atomistica_startup ...
// so that `import _atomistica` gives the reference's own extension types (Particles, Neighbors,
// Tersoff, ...) running on the GPU.  Built by atomistica_b200/seam2/build.py; the parameter
// defaults are generated from atomistica_b200/parameters.py into seam2_defaults.inc at build time.
#include <cmath>
#include <cstdarg>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <sstream>
#include <string>
#include <vector>

#include "atomistica_b200.h"
#include "ptrdict.h"   // the reference's header (section_t, ptrdict_register_*), -I src/support

#define ERROR_NONE 0
#define ERROR_UNSPECIFIED (-1)

// data type codes of src/python/c/atomisticamodule.h:39-49
enum { TYPE_REAL3x3_ATTR = 3, TYPE_INTEGER3_ATTR = 6, TYPE_INTEGER = 7, TYPE_REAL3 = 9 };

// ---------------------------------------------------------------------------------------------
// error stack (src/support/error.f90:177-289) and logging (f_logging.f90)
// ---------------------------------------------------------------------------------------------

struct ErrEntry {
  std::string fn, doc;
  int line, kind;
  bool has_doc;
};
static std::vector<ErrEntry> g_stack;
static FILE *g_log = nullptr;
static atx_ctx *g_ctx = nullptr;

static atx_ctx *ctx() {
  if (!g_ctx) {
    int dev = 0;
    if (const char *v = getenv("ATX_DEVICE")) dev = atoi(v);
    if (atx_ctx_create(dev, &g_ctx) != 0) g_ctx = nullptr;
  }
  return g_ctx;
}

extern "C" {

void c_push_error_with_info(const char *doc, const char *fn, int line, int kind) {
  g_stack.push_back({fn ? fn : "", doc ? doc : "", line, kind, true});
}
void c_push_error(const char *fn, int line, int kind) { g_stack.push_back({fn ? fn : "", "", line, kind, false}); }
void error_clear_stack(void) { g_stack.clear(); }
void c_error_abort(int kind) {
  fprintf(stderr, "Fatal error (kind %d) without an error variable to return it in.\n", kind);
  abort();
}
// "Traceback (most recent call last)" + one block per stack entry, newest first (error.f90:228-289)
void get_full_error_string(char *str) {
  std::string s = "Traceback (most recent call last)";
  for (size_t k = g_stack.size(); k-- > 0;) {
    const ErrEntry &e = g_stack[k];
    s += "\n  File \"" + e.fn + "\", line " + std::to_string(e.line);
    if (e.has_doc) s += "\n    " + e.doc;
  }
  g_stack.clear();
  if (s.size() > 9000) s.resize(9000);
  strcpy(str, s.c_str());
}

void f_logging_start(const char *fn) {
  if (g_log) fclose(g_log);
  g_log = fopen(fn, "w");
}
void c_prlog(const char *msg) {
  if (g_log) { fputs(msg, g_log); fputc('\n', g_log); }
}
void c_prscrlog(const char *msg) {
  c_prlog(msg);
  puts(msg);
}
void atomistica_startup(int) {
  if (!g_log) g_log = fopen("atomistica.log", "w");
  c_prlog("Atomistica (B200-native hot path behind the reference's Python module)");
  c_prlog(atx_version());
}
void atomistica_shutdown(void) {
  if (g_log) fclose(g_log);
  g_log = nullptr;
}

}  // extern "C"

// push the C ABI's message onto the stack; returns true when rc signals an error
static bool fail(int rc, int *ierror, const char *where, int line) {
  if (rc == 0) return false;
  char buf[2048];
  atx_last_error(buf, sizeof buf);
  c_push_error_with_info(buf, where, line, rc);
  if (ierror) *ierror = rc;
  else c_error_abort(rc);
  return true;
}
#define CHK(call, ierror)                                       \
  do {                                                          \
    if (fail((call), (ierror), __FILE__, __LINE__)) return;     \
  } while (0)
#define RAISE(ierror, msg)                                              \
  do {                                                                  \
    c_push_error_with_info((msg), __FILE__, __LINE__, ERROR_UNSPECIFIED); \
    if (ierror) { *(ierror) = ERROR_UNSPECIFIED; return; }              \
    c_error_abort(ERROR_UNSPECIFIED);                                   \
  } while (0)

// ---------------------------------------------------------------------------------------------
// particles_t (src/python/f90/python_particles.f90:84-170) and its data registry
// ---------------------------------------------------------------------------------------------

struct S2Particles {
  atx_particles *h = nullptr;
  int nat = 0;
  bool initialized = false;
  std::vector<int> Z, el;
  double *r = nullptr;          // (3,nat), page-locked: "coordinates"
  double cell[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1};   // Abox, column-major == ASE cell rows
  double bbox[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1};
  int pbc[3] = {1, 1, 1};
  std::vector<int> el2Z;
  void *tag = nullptr;
  bool pos_dirty = true;
};

static void particles_release(S2Particles *p) {
  if (p->r) atx_host_free_pinned(p->r);
  p->r = nullptr;
  p->Z.clear();
  p->el.clear();
  p->nat = 0;
  p->initialized = false;
}

static int particles_sync(S2Particles *p) {
  if (!p->pos_dirty) return 0;
  int rc = atx_particles_set_positions(p->h, p->nat, p->r);
  if (rc == 0) p->pos_dirty = false;
  return rc;
}

extern "C" {

void f_particles_new(void **out) {
  S2Particles *p = new S2Particles();
  if (ctx()) atx_particles_create(ctx(), &p->h);
  *out = p;
}
void f_particles_free(void *self) {
  S2Particles *p = (S2Particles *)self;
  particles_release(p);
  if (p->h) atx_particles_destroy(p->h);
  delete p;
}
void f_particles_init(void *) {}
void f_particles_del(void *self) { particles_release((S2Particles *)self); }

void f_particles_allocate(void *self, int nat, int *ierror) {
  S2Particles *p = (S2Particles *)self;
  if (!p->h) RAISE(ierror, "No CUDA device available: the particles object has no device side (no CPU fallback).");
  particles_release(p);
  p->nat = nat;
  p->Z.assign(nat, 0);
  p->el.assign(nat, 0);
  void *mem = nullptr;
  CHK(atx_host_alloc_pinned(sizeof(double) * 3 * (size_t)(nat > 0 ? nat : 1), &mem), ierror);
  p->r = (double *)mem;
  memset(p->r, 0, sizeof(double) * 3 * (size_t)nat);
  p->initialized = true;
  p->pos_dirty = true;
}

// particles_update_elements (python_particles.f90:617-658): compact ids by ascending Z
void f_particles_update_elements(void *self) {
  S2Particles *p = (S2Particles *)self;
  bool present[256] = {false};
  for (int z : p->Z)
    if (z > 0 && z < 256) present[z] = true;
  int z2el[256] = {0};
  p->el2Z.clear();
  for (int z = 1; z < 256; z++)
    if (present[z]) {
      p->el2Z.push_back(z);
      z2el[z] = (int)p->el2Z.size();
    }
  for (int i = 0; i < p->nat; i++) p->el[i] = (p->Z[i] > 0 && p->Z[i] < 256) ? z2el[p->Z[i]] : 0;
  if (p->h && p->nat > 0) fail(atx_particles_set_elements(p->h, p->nat, p->el.data()), nullptr, __FILE__, __LINE__);
}

// particles_set_cell (python_particles.f90:286-346): Abox columns = cell vectors, Bbox by gaussn
void f_particles_set_cell(void *self, double *cell, BOOL *pbc, int *ierror) {
  S2Particles *p = (S2Particles *)self;
  if (!p->h) RAISE(ierror, "No CUDA device available (no CPU fallback).");
  double A[9];
  for (int k = 0; k < 9; k++) { p->cell[k] = cell[k]; A[k] = cell[k]; p->bbox[k] = (k % 4 == 0) ? 1.0 : 0.0; }
  if (atx_host_gaussn(3, A, 3, p->bbox) != 0) RAISE(ierror, "Failed to determine the reciprocal lattice. Cell = singular?");
  for (int k = 0; k < 3; k++) p->pbc[k] = pbc[k] ? 1 : 0;
  CHK(atx_particles_set_cell(p->h, p->cell, p->bbox, p->pbc), ierror);
}

// particles_inbox (python_particles.f90:882-947): wrap into the cell along the periodic directions
void f_particles_inbox(void *self) {
  S2Particles *p = (S2Particles *)self;
  const double *A = p->cell, *B = p->bbox;
  for (int i = 0; i < p->nat; i++) {
    double *r = p->r + 3 * (size_t)i, s[3];
    for (int k = 0; k < 3; k++) {
      s[k] = B[k] * r[0] + B[3 + k] * r[1] + B[6 + k] * r[2];
      if (p->pbc[k]) s[k] -= std::floor(s[k]);
    }
    for (int c = 0; c < 3; c++) r[c] = A[c] * s[0] + A[3 + c] * s[1] + A[6 + c] * s[2];
  }
  p->pos_dirty = true;
}

void f_particles_i_changed_positions(void *self) { ((S2Particles *)self)->pos_dirty = true; }
void f_particles_get_data(void *self, void **data) { *data = self; }
void f_particles_set_tag(void *self, void *tag) { ((S2Particles *)self)->tag = tag; }
void f_particles_get_tag(void *self, void **tag) { *tag = ((S2Particles *)self)->tag; }
int f_particles_get_nel(void *self) { return (int)((S2Particles *)self)->el2Z.size(); }

// ---- data registry (src/python/f90/python_helper.f90:76-380, python_particles.f90:449-470)
int data_get_len(void *data) { return ((S2Particles *)data)->nat; }

BOOL f_data_exists(void *, const char *name, int *data_type) {
  struct { const char *n; int t; } known[] = {{"Z", TYPE_INTEGER}, {"atom_types", TYPE_INTEGER},
                                             {"internal_element_number", TYPE_INTEGER},
                                             {"coordinates", TYPE_REAL3}, {"cell", TYPE_REAL3x3_ATTR},
                                             {"pbc", TYPE_INTEGER3_ATTR}};
  for (auto &k : known)
    if (!strcmp(name, k.n)) { *data_type = k.t; return 1; }
  return 0;
}
void integer_ptr_by_name(void *data, const char *name, void **ptr, int *ierror) {
  S2Particles *p = (S2Particles *)data;
  if (!strcmp(name, "Z") || !strcmp(name, "atom_types")) *ptr = p->Z.data();
  else if (!strcmp(name, "internal_element_number")) *ptr = p->el.data();
  else RAISE(ierror, (std::string("Unknown integer field '") + name + "'.").c_str());
}
void realx_ptr_by_name(void *data, const char *name, void **ptr, int *ierror) {
  S2Particles *p = (S2Particles *)data;
  if (!strcmp(name, "coordinates")) *ptr = p->r;
  else RAISE(ierror, (std::string("Unknown real3 field '") + name + "'.").c_str());
}
void real_ptr_by_name(void *, const char *name, void **, int *ierror) {
  RAISE(ierror, (std::string("Unknown real field '") + name + "'.").c_str());
}
void realxxx_ptr_by_name(void *, const char *name, void **, int *ierror) {
  RAISE(ierror, (std::string("Unknown real3x3 field '") + name + "'.").c_str());
}
void real_attr_by_name(void *, const char *name, void **, int *ierror) {
  RAISE(ierror, (std::string("Unknown real attribute '") + name + "'.").c_str());
}
void real3_attr_by_name(void *, const char *name, void **, int *ierror) {
  RAISE(ierror, (std::string("Unknown real3 attribute '") + name + "'.").c_str());
}
void real3x3_attr_by_name(void *data, const char *name, void **ptr, int *ierror) {
  if (!strcmp(name, "cell")) *ptr = ((S2Particles *)data)->cell;
  else RAISE(ierror, (std::string("Unknown real3x3 attribute '") + name + "'.").c_str());
}
void integer_attr_by_name(void *, const char *name, void **, int *ierror) {
  RAISE(ierror, (std::string("Unknown integer attribute '") + name + "'.").c_str());
}
void integer3_attr_by_name(void *data, const char *name, void **ptr, int *ierror) {
  if (!strcmp(name, "pbc")) *ptr = ((S2Particles *)data)->pbc;
  else RAISE(ierror, (std::string("Unknown integer3 attribute '") + name + "'.").c_str());
}

}  // extern "C"

// ---------------------------------------------------------------------------------------------
// neighbors_t (src/python/f90/python_neighbors.f90:51-139) and its accessors (neighbors_wrap.f90)
// ---------------------------------------------------------------------------------------------

struct S2Neighbors {
  atx_neighbors *h = nullptr;
  int avgn = 100;
  double cutoff = 0.0, verlet_shell = 0.0;
  S2Particles *p = nullptr;
  void *tag = nullptr;
  // host copy of the list in the reference's layout, refreshed when the device list was rebuilt
  long long cached_build = -1;
  std::vector<intptr_t> seed, last;
  std::vector<int> nb, dc;
};

static void neighbors_ensure(S2Neighbors *n) {
  if (n->h || !ctx()) return;
  atx_neighbors_create(ctx(), n->avgn, &n->h);
  if (n->cutoff > 0.0) atx_neighbors_request_interaction_range(n->h, n->cutoff);
  if (n->verlet_shell > 0.0) atx_neighbors_set_verlet_shell(n->h, n->verlet_shell);
}

// seed/last/neighbors/dc on the host (f_get_* read them like neighbors_wrap.f90:211-550 does)
static int neighbors_host(S2Neighbors *n) {
  if (!n->h || !n->p) return ERROR_UNSPECIFIED;
  long long nbuilds = 0;
  atx_neighbors_get_counters(n->h, &nbuilds, nullptr);
  if (nbuilds == n->cached_build) return 0;
  long long npairs = 0;
  int rc = atx_neighbors_get_info(n->h, &npairs, nullptr, nullptr, nullptr);
  if (rc) return rc;
  const int nat = n->p->nat;
  const long long need = npairs + nat + 1;
  n->seed.assign(nat + 2, 0);
  n->last.assign(nat + 2, 0);
  n->nb.assign(need, 0);
  n->dc.assign(3 * need, 0);
  rc = atx_neighbors_copy_to_host(n->h, n->seed.data(), n->last.data(), n->nb.data(), n->dc.data(), need);
  if (rc) return rc;
  n->cached_build = nbuilds;
  return 0;
}

// dr = r_i - r_j + Abox.dc (macros.inc:76), 0-based atom i and 1-based list slot ni
static void bond_vector(const S2Neighbors *n, int i, intptr_t ni, double dr[3]) {
  const S2Particles *p = n->p;
  const int j = n->nb[ni - 1] - 1;
  const int *dc = &n->dc[3 * (ni - 1)];
  for (int c = 0; c < 3; c++)
    dr[c] = p->r[3 * (size_t)i + c] - p->r[3 * (size_t)j + c] + p->cell[c] * dc[0] + p->cell[3 + c] * dc[1] +
            p->cell[6 + c] * dc[2];
}

extern "C" {

void f_neighbors_new(void **out) { *out = new S2Neighbors(); }
void f_neighbors_free(void *self) {
  S2Neighbors *n = (S2Neighbors *)self;
  if (n->h) atx_neighbors_destroy(n->h);
  delete n;
}
void f_neighbors_init(void *self, int avgn) {
  S2Neighbors *n = (S2Neighbors *)self;
  n->avgn = avgn;
  neighbors_ensure(n);
}
void f_neighbors_del(void *self) {
  S2Neighbors *n = (S2Neighbors *)self;
  if (n->h) atx_neighbors_destroy(n->h);
  n->h = nullptr;
  n->cached_build = -1;
}
// neighbors_set (python_neighbors.f90:337-375); avgn / cutoff / verlet_shell < 0 mean "keep"
void f_neighbors_set(void *self, int avgn, double cutoff, double verlet_shell, double) {
  S2Neighbors *n = (S2Neighbors *)self;
  if (avgn > 0 && avgn != n->avgn) {
    n->avgn = avgn;
    if (n->h) { atx_neighbors_destroy(n->h); n->h = nullptr; n->cached_build = -1; }
  }
  neighbors_ensure(n);
  if (cutoff > 0.0) { n->cutoff = cutoff; if (n->h) atx_neighbors_request_interaction_range(n->h, cutoff); }
  if (verlet_shell >= 0.0) { n->verlet_shell = verlet_shell; if (n->h) atx_neighbors_set_verlet_shell(n->h, verlet_shell); }
}
void f_neighbors_request_interaction_range(void *self, double cutoff) {
  S2Neighbors *n = (S2Neighbors *)self;
  neighbors_ensure(n);
  if (cutoff > n->cutoff) n->cutoff = cutoff;
  if (n->h) atx_neighbors_request_interaction_range(n->h, cutoff);
}
void f_neighbors_update(void *self, void *pp, int *ierror) {
  S2Neighbors *n = (S2Neighbors *)self;
  S2Particles *p = (S2Particles *)pp;
  neighbors_ensure(n);
  if (!n->h) RAISE(ierror, "No CUDA device available (no CPU fallback).");
  if (ierror) *ierror = ERROR_NONE;
  n->p = p;
  CHK(particles_sync(p), ierror);
  CHK(atx_neighbors_update(n->h, p->h), ierror);
}
// neighbors_find_neighbor (python_neighbors.f90:969-999): 1-based atoms in, 1-based slots out (-1: none)
void f_neighbors_find_neighbor(void *self, int i, int j, int *n1, int *n2) {
  S2Neighbors *n = (S2Neighbors *)self;
  *n1 = *n2 = -1;
  if (neighbors_host(n)) return;
  for (intptr_t k = n->seed[i - 1]; k <= n->last[i - 1]; k++)
    if (n->nb[k - 1] == j) *n1 = (int)k;
  for (intptr_t k = n->seed[j - 1]; k <= n->last[j - 1]; k++)
    if (n->nb[k - 1] == i) *n2 = (int)k;
}
// neighbors_size = nat*avgn once the list exists (python_neighbors.f90:499), 0 before the first build
int f_get_neighbors_size(void *self) {
  S2Neighbors *n = (S2Neighbors *)self;
  long long nbuilds = 0;
  if (!n->h || !n->p) return 0;
  atx_neighbors_get_counters(n->h, &nbuilds, nullptr);
  return nbuilds > 0 ? n->p->nat * n->avgn : 0;
}
int f_get_coordination(void *self, int i, double cutoff) {   // 1-based atom (neighbors_wrap.f90:232)
  S2Neighbors *n = (S2Neighbors *)self;
  if (neighbors_host(n)) return 0;
  int c = 0;
  for (intptr_t k = n->seed[i - 1]; k <= n->last[i - 1]; k++) {
    double dr[3];
    bond_vector(n, i - 1, k, dr);
    if (dr[0] * dr[0] + dr[1] * dr[1] + dr[2] * dr[2] < cutoff * cutoff) c++;
  }
  return c;
}
int f_get_coordination_numbers(void *self, double cutoff, int *c) {
  S2Neighbors *n = (S2Neighbors *)self;
  if (neighbors_host(n)) return 0;
  for (int i = 0; i < n->p->nat; i++) {
    c[i] = 0;
    for (intptr_t k = n->seed[i]; k <= n->last[i]; k++) {
      double dr[3];
      bond_vector(n, i, k, dr);
      if (dr[0] * dr[0] + dr[1] * dr[1] + dr[2] * dr[2] < cutoff * cutoff) c[i]++;
    }
  }
  return 0;
}
int f_get_number_of_neighbors(void *self, int i) {   // 0-based atom (neighbors_wrap.f90:308)
  S2Neighbors *n = (S2Neighbors *)self;
  if (neighbors_host(n)) return 0;
  return (int)(n->last[i] - n->seed[i] + 1);
}
int f_get_number_of_all_neighbors(void *self) {
  S2Neighbors *n = (S2Neighbors *)self;
  long long npairs = 0;
  if (!n->h) return 0;
  atx_neighbors_get_info(n->h, &npairs, nullptr, nullptr, nullptr);
  return (int)npairs;
}
void f_get_neighbors(void *self, int i, int *i2, double *r) {
  S2Neighbors *n = (S2Neighbors *)self;
  if (neighbors_host(n)) return;
  int q = 0;
  for (intptr_t k = n->seed[i]; k <= n->last[i]; k++, q++) {
    double dr[3];
    bond_vector(n, i, k, dr);
    i2[q] = n->nb[k - 1] - 1;
    r[q] = std::sqrt(dr[0] * dr[0] + dr[1] * dr[1] + dr[2] * dr[2]);
  }
}
void f_get_seed(void *self, int *seed) {
  S2Neighbors *n = (S2Neighbors *)self;
  if (neighbors_host(n)) return;
  seed[0] = 0;
  for (int i = 1; i < n->p->nat; i++) seed[i] = seed[i - 1] + (int)(n->last[i - 1] - n->seed[i - 1] + 1);
}
void f_get_all_neighbors(void *self, int *i1, int *i2, double *r) {
  S2Neighbors *n = (S2Neighbors *)self;
  if (neighbors_host(n)) return;
  size_t q = 0;
  for (int i = 0; i < n->p->nat; i++)
    for (intptr_t k = n->seed[i]; k <= n->last[i]; k++, q++) {
      double dr[3];
      bond_vector(n, i, k, dr);
      i1[q] = i;
      i2[q] = n->nb[k - 1] - 1;
      r[q] = std::sqrt(dr[0] * dr[0] + dr[1] * dr[1] + dr[2] * dr[2]);
    }
}
void f_get_all_neighbors_vec(void *self, int *i1, int *i2, double *drv, double *abs_dr) {
  S2Neighbors *n = (S2Neighbors *)self;
  if (neighbors_host(n)) return;
  size_t q = 0;
  for (int i = 0; i < n->p->nat; i++)
    for (intptr_t k = n->seed[i]; k <= n->last[i]; k++, q++) {
      double dr[3];
      bond_vector(n, i, k, dr);
      i1[q] = i;
      i2[q] = n->nb[k - 1] - 1;
      for (int c = 0; c < 3; c++) drv[3 * q + c] = dr[c];
      abs_dr[q] = std::sqrt(dr[0] * dr[0] + dr[1] * dr[1] + dr[2] * dr[2]);
    }
}
// list-slot order -> dense order (python_neighbors.f90:1030-1051, neighbors_wrap.f90:520-547)
void f_pack_per_bond_scalar(void *self, double *r1, double *r2) {
  S2Neighbors *n = (S2Neighbors *)self;
  if (neighbors_host(n)) return;
  size_t q = 0;
  for (int i = 0; i < n->p->nat; i++)
    for (intptr_t k = n->seed[i]; k <= n->last[i]; k++, q++) r2[q] = r1[k - 1];
}
void f_pack_per_bond_3x3(void *self, double *r1, double *r2) {
  S2Neighbors *n = (S2Neighbors *)self;
  if (neighbors_host(n)) return;
  size_t q = 0;
  for (int i = 0; i < n->p->nat; i++)
    for (intptr_t k = n->seed[i]; k <= n->last[i]; k++, q++) memcpy(r2 + 9 * q, r1 + 9 * (size_t)(k - 1), 9 * sizeof(double));
}
void f_neighbors_set_tag(void *self, void *tag) { ((S2Neighbors *)self)->tag = tag; }
void f_neighbors_get_tag(void *self, void **tag) { *tag = ((S2Neighbors *)self)->tag; }

}  // extern "C"

// ---------------------------------------------------------------------------------------------
// potential classes (what src/python/gen_factory.py:83-360 generates around each Fortran module)
// ---------------------------------------------------------------------------------------------

static const char *SYMBOLS[] = {"X", "H", "He", "Li", "Be", "B", "C", "N", "O", "F", "Ne", "Na", "Mg", "Al", "Si", "P",
                                "S", "Cl", "Ar", "K", "Ca", "Sc", "Ti", "V", "Cr", "Mn", "Fe", "Co", "Ni", "Cu", "Zn",
                                "Ga", "Ge", "As", "Se", "Br", "Kr", "Rb", "Sr", "Y", "Zr", "Nb", "Mo", "Tc", "Ru", "Rh",
                                "Pd", "Ag", "Cd", "In", "Sn", "Sb", "Te", "I", "Xe", "Cs", "Ba", "La", "Ce", "Pr", "Nd",
                                "Pm", "Sm", "Eu", "Gd", "Tb", "Dy", "Ho", "Er", "Tm", "Yb", "Lu", "Hf", "Ta", "W", "Re",
                                "Os", "Ir", "Pt", "Au", "Hg", "Tl", "Pb", "Bi"};
static int symbol_to_Z(const char *s2) {   // two characters, blank padded (Fortran string)
  char s[3] = {s2[0], (s2[1] == ' ' || s2[1] == 0) ? (char)0 : s2[1], 0};
  for (int z = 1; z < (int)(sizeof(SYMBOLS) / sizeof(SYMBOLS[0])); z++)
    if (!strcmp(SYMBOLS[z], s)) return z;
  return 0;
}

// Who frees the ptrdict section new_instance hands out?  In the Python flavour nobody but free_instance can
// (potential_dealloc, src/python/c/potential.c:93-101, only calls free_instance); the LAMMPS pair style cleans
// `members_` itself before free_instance (pair_atomistica.cpp:137-142, :296-306), so there the instance must
// leave it alone -- ptrdict_cleanup frees the root (c_ptrdict.c:380-402) and a second call is a double free.
#ifdef ATX_SEAM2_LAMMPS
#define S2_CLEANUP_MEMBERS(m) ((void)(m))
#else
#define S2_CLEANUP_MEMBERS(m) do { if (m) ptrdict_cleanup(m); } while (0)
#endif

#define S2_EL ATX_BOP_MAX_EL
#define S2_PAIRS ATX_BOP_MAX_PAIRS

// Bond-order instance: the parameter database in the layout of the Fortran BOP_DB_TYPE, one length
// counter per list property (the Fortran types carry nA, nB, ... the same way)
struct S2Bop {
  int kind = 0;          // ATX_BOP_*
  bool screened = false;
  char el[S2_EL][2];
  int nel = 0;
  char ref[128];
  atx_bop_params par;
  atx_bop_screening scr;
  int len[40];
  atx_bop *h = nullptr;
  section_t *members = nullptr;
};

// Juslin / JuslinScr instance (classes further down)
struct S2Juslin {
  bool screened = false;
  char el[3][2];
  int nel = 0;
  char ref[128];
  atx_juslin_params par;
  atx_juslin_screening scr;
  int len[24];
  atx_bop *h = nullptr;
  section_t *members = nullptr;
};

#include "seam2_defaults.inc"   // static void s2_bop_defaults(S2Bop *b): generated from parameters.py

static void reg_list(section_t *m, S2Bop *b, int &k, double *ptr, int maxlen, int n, const char *name) {
  b->len[k] = n;
  ptrdict_register_list_property(m, ptr, maxlen, &b->len[k], (char *)name, (char *)"See functional form.");
  k++;
}
static void reg_ilist(section_t *m, S2Bop *b, int &k, int *ptr, int maxlen, int n, const char *name) {
  b->len[k] = n;
  ptrdict_register_integer_list_property(m, (double *)ptr, maxlen, &b->len[k], (char *)name,
                                         (char *)"See functional form.");
  k++;
}

// REGISTER_FUNC of tersoff_registry.f90:22-116, kumagai_registry.f90, brenner_registry.f90
static void bop_register(S2Bop *b, section_t *cfg, section_t **members, const char *name, const char *descr) {
  s2_bop_defaults(b);
  section_t *m = ptrdict_register_section(cfg, (char *)name, (char *)descr);
  const int ne = b->nel, np = ne * (ne + 1) / 2;
  ptrdict_register_string_list_property(m, &b->el[0][0], 2, S2_EL, &b->nel, (char *)"el",
                                        (char *)"List of element symbols.");
  atx_bop_params &p = b->par;
  int k = 0;
  if (b->kind == ATX_BOP_TERSOFF) {
    reg_list(m, b, k, p.A, S2_PAIRS, np, "A"); reg_list(m, b, k, p.B, S2_PAIRS, np, "B");
    reg_list(m, b, k, p.xi, S2_PAIRS, np, "xi"); reg_list(m, b, k, p.lambda, S2_PAIRS, np, "lambda");
    reg_list(m, b, k, p.mu, S2_PAIRS, np, "mu"); reg_list(m, b, k, p.omega, S2_PAIRS, np, "omega");
    reg_list(m, b, k, p.mubo, S2_PAIRS, np, "mubo"); reg_ilist(m, b, k, p.m, S2_PAIRS, np, "m");
    reg_list(m, b, k, p.beta, S2_EL, ne, "beta"); reg_list(m, b, k, p.n, S2_EL, ne, "n");
    reg_list(m, b, k, p.c, S2_EL, ne, "c"); reg_list(m, b, k, p.d, S2_EL, ne, "d");
    reg_list(m, b, k, p.h, S2_EL, ne, "h");
  } else if (b->kind == ATX_BOP_KUMAGAI) {
    reg_list(m, b, k, p.A, S2_PAIRS, np, "A"); reg_list(m, b, k, p.B, S2_PAIRS, np, "B");
    reg_list(m, b, k, p.lambda, S2_PAIRS, np, "lambda1"); reg_list(m, b, k, p.mu, S2_PAIRS, np, "lambda2");
    reg_list(m, b, k, p.eta, S2_EL, ne, "eta"); reg_list(m, b, k, p.delta, S2_EL, ne, "delta");
    reg_list(m, b, k, p.mubo, S2_PAIRS, np, "alpha"); reg_ilist(m, b, k, p.m, S2_PAIRS, np, "beta");
    reg_list(m, b, k, p.c1, S2_EL, ne, "c1"); reg_list(m, b, k, p.c2, S2_EL, ne, "c2");
    reg_list(m, b, k, p.c3, S2_EL, ne, "c3"); reg_list(m, b, k, p.c4, S2_EL, ne, "c4");
    reg_list(m, b, k, p.c5, S2_EL, ne, "c5"); reg_list(m, b, k, p.h, S2_EL, ne, "h");
  } else {
    memset(b->ref, ' ', sizeof b->ref);
    ptrdict_register_string_property(m, b->ref, (int)sizeof b->ref, (char *)"ref",
                                     (char *)"Reference string to choose a parameters set from the database.");
    reg_list(m, b, k, p.D0, S2_PAIRS, np, "D0"); reg_list(m, b, k, p.r0, S2_PAIRS, np, "r0");
    reg_list(m, b, k, p.S, S2_PAIRS, np, "S"); reg_list(m, b, k, p.pbeta, S2_PAIRS, np, "beta");
    reg_list(m, b, k, p.gamma, S2_PAIRS, np, "gamma"); reg_list(m, b, k, p.pc, S2_PAIRS, np, "c");
    reg_list(m, b, k, p.pd, S2_PAIRS, np, "d"); reg_list(m, b, k, p.ph, S2_PAIRS, np, "h");
    reg_list(m, b, k, p.mubo, S2_PAIRS, np, "mu"); reg_list(m, b, k, p.pn, S2_PAIRS, np, "n");
    reg_ilist(m, b, k, p.m, S2_PAIRS, np, "m");
  }
  reg_list(m, b, k, p.r1, S2_PAIRS, np, "r1"); reg_list(m, b, k, p.r2, S2_PAIRS, np, "r2");
  if (b->screened) {
    reg_list(m, b, k, b->scr.or1, S2_PAIRS, np, "or1"); reg_list(m, b, k, b->scr.or2, S2_PAIRS, np, "or2");
    reg_list(m, b, k, b->scr.bor1, S2_PAIRS, np, "bor1"); reg_list(m, b, k, b->scr.bor2, S2_PAIRS, np, "bor2");
    reg_list(m, b, k, b->scr.Cmin, S2_PAIRS, np, "Cmin"); reg_list(m, b, k, b->scr.Cmax, S2_PAIRS, np, "Cmax");
  }
  b->members = m;
  *members = m;
}

static void bop_free(void *self) {
  S2Bop *b = (S2Bop *)self;
  if (b->h) atx_bop_destroy(b->h);
  S2_CLEANUP_MEMBERS(b->members);
  delete b;
}
static void bop_register_data(void *, void *, int *ierror) { if (ierror) *ierror = ERROR_NONE; }

// INIT_FUNC: the keyword arguments have been written through the registered pointers; (re)create
// the device object from the database
static void bop_init(void *self, int *ierror) {
  S2Bop *b = (S2Bop *)self;
  if (ierror) *ierror = ERROR_NONE;
  if (!ctx()) RAISE(ierror, "No CUDA device available (no CPU fallback).");
  if (b->nel < 1 || b->nel > S2_EL) RAISE(ierror, "Number of elements out of range.");
  b->par.kind = b->kind;
  b->par.nel = b->nel;
  for (int i = 0; i < b->nel; i++) {
    b->par.Z[i] = symbol_to_Z(b->el[i]);
    if (b->par.Z[i] <= 0) RAISE(ierror, "Unknown element symbol in 'el'.");
  }
  if (b->h) { atx_bop_destroy(b->h); b->h = nullptr; }
  if (b->screened) CHK(atx_bop_create_screened(ctx(), &b->par, &b->scr, &b->h), ierror);
  else CHK(atx_bop_create(ctx(), &b->par, &b->h), ierror);
}

// BIND_TO_FUNC (default_bind_to_func.f90:25-146): element map, interaction range request
static void bop_bind_to(void *self, void *pp, void *nn, int *ierror) {
  S2Bop *b = (S2Bop *)self;
  S2Particles *p = (S2Particles *)pp;
  S2Neighbors *n = (S2Neighbors *)nn;
  if (ierror) *ierror = ERROR_NONE;
  if (!b->h) RAISE(ierror, "The potential has not been initialised.");
  neighbors_ensure(n);
  if (!n->h || !p->h) RAISE(ierror, "No CUDA device available (no CPU fallback).");
  n->p = p;
  CHK(atx_bop_bind_to(b->h, p->h, n->h, (int)p->el2Z.size(), p->el2Z.data()), ierror);
}

// COMPUTE_FUNC (default_compute_func.f90:25-102); argument list of gen_factory.py:239-356
// (the template's prototype types wpot as int* and mask as double*, src/python/c/factory.template.h:
// the Fortran side is authoritative -- wpot is real(3,3), mask integer(nat))
static void bop_energy_and_forces(void *self, void *pp, void *nn, double *, double *epot, double *f, int *wpot_,
                                  double *mask_, double *epot_per_at, double *epot_per_bond, double *f_per_bond,
                                  double *wpot_per_at, double *wpot_per_bond, int *ierror) {
  double *wpot = (double *)wpot_;
  int *mask = (int *)mask_;
  S2Bop *b = (S2Bop *)self;
  S2Particles *p = (S2Particles *)pp;
  S2Neighbors *n = (S2Neighbors *)nn;
  if (ierror) *ierror = ERROR_NONE;
  if (!b->h || !n->h) RAISE(ierror, "bind_to has not been called on this potential.");
  n->p = p;
  CHK(particles_sync(p), ierror);
  CHK(atx_bop_energy_and_forces(b->h, p->h, n->h, mask, epot, f, wpot, epot_per_at, epot_per_bond, f_per_bond,
                                wpot_per_at, wpot_per_bond), ierror);
}

template <int KIND, bool SCR>
static void bop_new(void **self, section_t *cfg, section_t **members) {
  S2Bop *b = new S2Bop();
  memset(&b->par, 0, sizeof b->par);
  memset(&b->scr, 0, sizeof b->scr);
  b->kind = KIND;
  b->screened = SCR;
  static const char *names[] = {"", "Tersoff", "Kumagai", "Brenner"};
  std::string nm = std::string(names[KIND]) + (SCR ? "Scr" : "");
  bop_register(b, cfg, members, nm.c_str(), "Bond-order potential (B200-native kernels, csrc/atx_bop.cu).");
  *self = b;
}

// ---- TabulatedAlloyEAM (tabulated_alloy_eam.f90:147-259 init, :297-350 bind_to, :360-415 compute)

struct S2Spline {
  std::vector<double> y, d2y, c1, c2, c3, d1, d2, d3;
  atx_spline s;
};
struct S2Eam {
  char elements[1024];
  char fn[100];
  BOOL dump = 0;
  std::vector<std::string> names;
  std::vector<S2Spline *> keep;
  double cutoff = 0.0;
  atx_eam *h = nullptr;
  section_t *members = nullptr;
};

static std::string fstr(const char *s, int n) {
  std::string r(s, strnlen(s, n));
  while (!r.empty() && (r.back() == ' ' || r.back() == 0)) r.pop_back();
  return r;
}

static S2Spline *make_spline(int n, double dx, const std::vector<double> &y_in, double scale) {
  S2Spline *sp = new S2Spline();
  sp->y.assign(n, 0); sp->d2y.assign(n, 0);
  for (auto *v : {&sp->c1, &sp->c2, &sp->c3, &sp->d1, &sp->d2, &sp->d3}) v->assign(n - 1, 0);
  atx_host_spline_init(n, 0.0, dx, y_in.data(), sp->y.data(), sp->d2y.data(), sp->c1.data(), sp->c2.data(),
                       sp->c3.data(), sp->d1.data(), sp->d2.data(), sp->d3.data());
  if (scale != 1.0)   // simple_spline_scale_y_axis
    for (auto *v : {&sp->y, &sp->d2y, &sp->c1, &sp->c2, &sp->c3, &sp->d1, &sp->d2, &sp->d3})
      for (double &x : *v) x *= scale;
  sp->s.n = n; sp->s.x0 = 0.0; sp->s.dx = dx;
  sp->s.y = sp->y.data(); sp->s.coeff1 = sp->c1.data(); sp->s.coeff2 = sp->c2.data(); sp->s.coeff3 = sp->c3.data();
  sp->s.dcoeff1 = sp->d1.data(); sp->s.dcoeff2 = sp->d2.data(); sp->s.dcoeff3 = sp->d3.data();
  return sp;
}

static void eam_new(void **self, section_t *cfg, section_t **members) {
  S2Eam *e = new S2Eam();
  memset(e->elements, ' ', sizeof e->elements);
  e->elements[0] = '*';
  memset(e->fn, ' ', sizeof e->fn);
  memcpy(e->fn, "default.in", 10);
  section_t *m = ptrdict_register_section(cfg, (char *)"TabulatedAlloyEAM",
                                          (char *)"General tabulated EAM potential for many component systems (alloys).");
  ptrdict_register_string_property(m, e->elements, (int)sizeof e->elements, (char *)"elements",
                                   (char *)"Element for which to use this potential.");
  ptrdict_register_string_property(m, e->fn, (int)sizeof e->fn, (char *)"fn", (char *)"Configuration file.");
  ptrdict_register_boolean_property(m, &e->dump, (char *)"dump", (char *)"Dump interatomic potential to disk.");
  e->members = m;
  *members = m;
  *self = e;
}
static void eam_free(void *self) {
  S2Eam *e = (S2Eam *)self;
  if (e->h) atx_eam_destroy(e->h);
  for (auto *s : e->keep) delete s;
  S2_CLEANUP_MEMBERS(e->members);
  delete e;
}
// setfl reader: 3 comment lines; nel names; nF dF nr dr cutoff; per element header + F + rho; then
// the lower triangle of r*phi tables.  rho and r*phi are padded with two zeros, r*phi scaled by 1/2.
static void eam_init(void *self, int *ierror) {
  S2Eam *e = (S2Eam *)self;
  if (ierror) *ierror = ERROR_NONE;
  if (!ctx()) RAISE(ierror, "No CUDA device available (no CPU fallback).");
  const std::string fn = fstr(e->fn, sizeof e->fn);
  std::ifstream in(fn);
  if (!in) RAISE(ierror, ("Error opening file '" + fn + "'.").c_str());
  std::string line;
  for (int k = 0; k < 3; k++) std::getline(in, line);
  std::getline(in, line);
  std::istringstream hs(line);
  int nel = 0;
  hs >> nel;
  if (nel < 1 || nel > 10) RAISE(ierror, "Number of elements in the setfl file out of range.");
  e->names.assign(nel, "");
  for (auto &s : e->names) hs >> s;
  std::string rest((std::istreambuf_iterator<char>(in)), std::istreambuf_iterator<char>());
  for (char &c : rest)
    if (c == 'D' || c == 'd') c = 'E';
  std::istringstream ts(rest);
  int nF = 0, nr = 0;
  double dF = 0, dr = 0;
  ts >> nF >> dF >> nr >> dr >> e->cutoff;
  if (!ts || nF < 2 || nr < 2) RAISE(ierror, "Malformed setfl header.");
  std::vector<atx_spline> fF(nel), frho(nel), fphi((size_t)nel * nel);
  for (int i = 0; i < nel; i++) {
    double zi, mass, a0;
    std::string lat;
    ts >> zi >> mass >> a0 >> lat;
    std::vector<double> F(nF), rho(nr + 2, 0.0);
    for (double &x : F) ts >> x;
    for (int k = 0; k < nr; k++) ts >> rho[k];
    S2Spline *a = make_spline(nF, dF, F, 1.0), *b = make_spline(nr + 2, dr, rho, 1.0);
    e->keep.push_back(a); e->keep.push_back(b);
    fF[i] = a->s; frho[i] = b->s;
  }
  for (int i = 0; i < nel; i++)
    for (int j = 0; j <= i; j++) {
      std::vector<double> rphi(nr + 2, 0.0);
      for (int k = 0; k < nr; k++) ts >> rphi[k];
      S2Spline *a = make_spline(nr + 2, dr, rphi, 0.5);
      e->keep.push_back(a);
      fphi[(size_t)i + (size_t)nel * j] = a->s;
      fphi[(size_t)j + (size_t)nel * i] = a->s;
    }
  if (!ts) RAISE(ierror, ("Unexpected end of file in '" + fn + "'.").c_str());
  if (e->h) { atx_eam_destroy(e->h); e->h = nullptr; }
  CHK(atx_eam_create(ctx(), nel, fF.data(), frho.data(), fphi.data(), e->cutoff, &e->h), ierror);
}
static void eam_bind_to(void *self, void *pp, void *nn, int *ierror) {
  S2Eam *e = (S2Eam *)self;
  S2Particles *p = (S2Particles *)pp;
  S2Neighbors *n = (S2Neighbors *)nn;
  if (ierror) *ierror = ERROR_NONE;
  if (!e->h) RAISE(ierror, "The potential has not been initialised.");
  neighbors_ensure(n);
  if (!n->h || !p->h) RAISE(ierror, "No CUDA device available (no CPU fallback).");
  n->p = p;
  std::vector<int> el2db;
  for (int z : p->el2Z) {
    int db = -1;
    for (size_t k = 0; k < e->names.size(); k++)
      if (z > 0 && z < (int)(sizeof(SYMBOLS) / sizeof(SYMBOLS[0])) && e->names[k] == SYMBOLS[z]) db = (int)k + 1;
    el2db.push_back(db);
  }
  CHK(atx_eam_bind_to(e->h, p->h, n->h, (int)el2db.size(), el2db.data()), ierror);
}
static void eam_energy_and_forces(void *self, void *pp, void *nn, double *, double *epot, double *f, int *wpot_,
                                  double *mask_, double *epot_per_at, double *epot_per_bond, double *f_per_bond,
                                  double *wpot_per_at, double *wpot_per_bond, int *ierror) {
  double *wpot = (double *)wpot_;
  int *mask = (int *)mask_;
  S2Eam *e = (S2Eam *)self;
  S2Particles *p = (S2Particles *)pp;
  S2Neighbors *n = (S2Neighbors *)nn;
  if (ierror) *ierror = ERROR_NONE;
  if (epot_per_bond || f_per_bond || wpot_per_bond)
    RAISE(ierror, "TabulatedAlloyEAM does not support per-bond properties.");   // features: mask, per_at
  if (!e->h || !n->h) RAISE(ierror, "bind_to has not been called on this potential.");
  n->p = p;
  CHK(particles_sync(p), ierror);
  CHK(atx_eam_energy_and_forces(e->h, p->h, n->h, mask, epot, f, wpot, epot_per_at, wpot_per_at), ierror);
}
static void eam_register_data(void *, void *, int *ierror) { if (ierror) *ierror = ERROR_NONE; }


// ---- Juslin / JuslinScr (juslin_registry.f90:22-125, juslin_module.f90:209-420) -----------------------
// nel**2 pair rows (PAIR_INDEX_NS) and nel**3 triplet rows; the defaults are the W-C-H set of
// juslin_params.f90 (generated from parameters.py); INIT mirrors the rows with r0 < 0 from the transposed
// pair (BIND_TO_FUNC :283-312) before the device object is created.  `ref` is registered like the
// reference does; the database it would select from has a single entry per class here.


template <bool SCR>
static void juslin_new(void **self, section_t *cfg, section_t **members) {
  S2Juslin *b = new S2Juslin();
  memset(&b->par, 0, sizeof b->par);
  memset(&b->scr, 0, sizeof b->scr);
  b->screened = SCR;
  s2_juslin_defaults(b);
  section_t *m = ptrdict_register_section(cfg, (char *)(SCR ? "JuslinScr" : "Juslin"),
                                          (char *)(SCR ? "Juslin-Type bond-order potential (screened)."
                                                       : "Juslin-Type bond-order potential."));
  ptrdict_register_string_list_property(m, &b->el[0][0], 2, 3, &b->nel, (char *)"el", (char *)"List of element symbols.");
  memset(b->ref, ' ', sizeof b->ref);
  ptrdict_register_string_property(m, b->ref, (int)sizeof b->ref, (char *)"ref",
                                   (char *)"Reference string to choose a parameters set from the database.");
  const int np = b->nel * b->nel, nt = np * b->nel;
  int k = 0;
  auto reg = [&](double *ptr, int maxlen, int n, const char *name) {
    b->len[k] = n;
    ptrdict_register_list_property(m, ptr, maxlen, &b->len[k], (char *)name, (char *)"See functional form.");
    k++;
  };
  atx_juslin_params &q = b->par;
  reg(q.D0, 9, np, "D0"); reg(q.r0, 9, np, "r0"); reg(q.S, 9, np, "S"); reg(q.beta, 9, np, "beta");
  reg(q.gamma, 9, np, "gamma"); reg(q.c, 9, np, "c"); reg(q.d, 9, np, "d"); reg(q.h, 9, np, "h");
  reg(q.n, 9, np, "n"); reg(q.alpha, 27, nt, "alpha"); reg(q.omega, 27, nt, "omega");
  b->len[k] = nt;
  ptrdict_register_integer_list_property(m, (double *)q.m, 27, &b->len[k], (char *)"m", (char *)"See functional form.");
  k++;
  reg(q.r1, 9, np, "r1"); reg(q.r2, 9, np, "r2");
  if (SCR) {
    reg(b->scr.or1, 9, np, "or1"); reg(b->scr.or2, 9, np, "or2"); reg(b->scr.bor1, 9, np, "bor1");
    reg(b->scr.bor2, 9, np, "bor2"); reg(b->scr.Cmin, 9, np, "Cmin"); reg(b->scr.Cmax, 9, np, "Cmax");
  }
  b->members = m;
  *members = m;
  *self = b;
}
static void juslin_free(void *self) {
  S2Juslin *b = (S2Juslin *)self;
  if (b->h) atx_bop_destroy(b->h);
  S2_CLEANUP_MEMBERS(b->members);
  delete b;
}
static void juslin_init(void *self, int *ierror) {
  S2Juslin *b = (S2Juslin *)self;
  if (ierror) *ierror = ERROR_NONE;
  if (!ctx()) RAISE(ierror, "No CUDA device available (no CPU fallback).");
  if (b->nel < 1 || b->nel > 3) RAISE(ierror, "Number of elements out of range.");
  atx_juslin_params q = b->par;          // the registered database stays as the host wrote it
  atx_juslin_screening sc = b->scr;
  q.nel = b->nel;
  for (int i = 0; i < b->nel; i++) {
    q.Z[i] = symbol_to_Z(b->el[i]);
    if (q.Z[i] <= 0) RAISE(ierror, "Unknown element symbol in 'el'.");
  }
  const int ne = b->nel;
  for (int i = 0; i < ne; i++)
    for (int j = 0; j < ne; j++) {
      const int a = j + i * ne, t = i + j * ne;   // PAIR_INDEX_NS(i,j), PAIR_INDEX_NS(j,i), 0-based
      if (q.r0[a] < 0.0) {
        double *rows[] = {q.D0, q.r0, q.S, q.beta, q.gamma, q.c, q.d, q.h, q.n, q.r1, q.r2,
                          sc.or1, sc.or2, sc.bor1, sc.bor2, sc.Cmin, sc.Cmax};
        for (double *row : rows) row[a] = row[t];
      }
    }
  if (b->h) { atx_bop_destroy(b->h); b->h = nullptr; }
  if (b->screened) CHK(atx_bop_create_juslin_screened(ctx(), &q, &sc, &b->h), ierror);
  else CHK(atx_bop_create_juslin(ctx(), &q, &b->h), ierror);
}
static void juslin_bind_to(void *self, void *pp, void *nn, int *ierror) {
  S2Juslin *b = (S2Juslin *)self;
  S2Particles *p = (S2Particles *)pp;
  S2Neighbors *n = (S2Neighbors *)nn;
  if (ierror) *ierror = ERROR_NONE;
  if (!b->h) RAISE(ierror, "The potential has not been initialised.");
  neighbors_ensure(n);
  if (!n->h || !p->h) RAISE(ierror, "No CUDA device available (no CPU fallback).");
  n->p = p;
  CHK(atx_bop_bind_to(b->h, p->h, n->h, (int)p->el2Z.size(), p->el2Z.data()), ierror);
}
static void juslin_energy_and_forces(void *self, void *pp, void *nn, double *, double *epot, double *f, int *wpot_,
                                     double *mask_, double *epot_per_at, double *epot_per_bond, double *f_per_bond,
                                     double *wpot_per_at, double *wpot_per_bond, int *ierror) {
  S2Juslin *b = (S2Juslin *)self;
  S2Particles *p = (S2Particles *)pp;
  S2Neighbors *n = (S2Neighbors *)nn;
  if (ierror) *ierror = ERROR_NONE;
  if (!b->h || !n->h) RAISE(ierror, "bind_to has not been called on this potential.");
  n->p = p;
  CHK(particles_sync(p), ierror);
  CHK(atx_bop_energy_and_forces(b->h, p->h, n->h, (int *)mask_, epot, f, (double *)wpot_, epot_per_at, epot_per_bond,
                                f_per_bond, wpot_per_at, wpot_per_bond), ierror);
}

// ---- TabulatedEAM (tabulated_eam.f90:141-214 init with the funcfl reader, :238-261 bind_to, :277-330) --
struct S2Funcfl {
  char elements[1024];
  char fn[100];
  std::vector<S2Spline *> keep;
  double cutoff = 0.0;
  atx_eam *h = nullptr;
  section_t *members = nullptr;
};
static void funcfl_new(void **self, section_t *cfg, section_t **members) {
  S2Funcfl *e = new S2Funcfl();
  memset(e->elements, ' ', sizeof e->elements);
  e->elements[0] = '*';
  memset(e->fn, ' ', sizeof e->fn);
  memcpy(e->fn, "default.in", 10);
  section_t *m = ptrdict_register_section(
      cfg, (char *)"TabulatedEAM",
      (char *)"General tabulated EAM potential, see S.M. Foiles, M.I. Baskes, M.S. Daw, Phys. Rev. B 33, 7983 (1986).");
  ptrdict_register_string_property(m, e->elements, (int)sizeof e->elements, (char *)"elements",
                                   (char *)"Element for which to use this potential.");
  ptrdict_register_string_property(m, e->fn, (int)sizeof e->fn, (char *)"fn", (char *)"Configuration file.");
  e->members = m;
  *members = m;
  *self = e;
}
static void funcfl_free(void *self) {
  S2Funcfl *e = (S2Funcfl *)self;
  if (e->h) atx_eam_destroy(e->h);
  for (auto *s : e->keep) delete s;
  S2_CLEANUP_MEMBERS(e->members);
  delete e;
}
// funcfl reader (:172-197): comment; Z mass a0 lattice; nF dF nr dr cutoff; nF F values, nr Z values,
// nr rho values.  Z is scaled by sqrt(0.5 Hartree Bohr) (:199; Units.f90:74-76).
static void funcfl_init(void *self, int *ierror) {
  S2Funcfl *e = (S2Funcfl *)self;
  if (ierror) *ierror = ERROR_NONE;
  if (!ctx()) RAISE(ierror, "No CUDA device available (no CPU fallback).");
  const std::string fn = fstr(e->fn, sizeof e->fn);
  std::ifstream in(fn);
  if (!in) RAISE(ierror, ("Error opening file '" + fn + "'.").c_str());
  std::string line;
  std::getline(in, line);
  std::getline(in, line);   // Z mass a0 lattice: not used by the kernel
  std::string rest((std::istreambuf_iterator<char>(in)), std::istreambuf_iterator<char>());
  for (char &c : rest)
    if (c == 'D' || c == 'd') c = 'E';
  std::istringstream ts(rest);
  int nF = 0, nr = 0;
  double dF = 0, dr = 0;
  ts >> nF >> dF >> nr >> dr >> e->cutoff;
  if (!ts || nF < 2 || nr < 2) RAISE(ierror, "Malformed funcfl header.");
  std::vector<double> F(nF), Z(nr), rho(nr);
  for (double &x : F) ts >> x;
  for (double &x : Z) ts >> x;
  for (double &x : rho) ts >> x;
  if (!ts) RAISE(ierror, ("Unexpected end of file in '" + fn + "'.").c_str());
  const double HARTREE = 27.2113961, BOHR = 0.529177249;
  S2Spline *sF = make_spline(nF, dF, F, 1.0), *sZ = make_spline(nr, dr, Z, std::sqrt(0.5 * HARTREE * BOHR)),
           *sr = make_spline(nr, dr, rho, 1.0);
  e->keep.push_back(sF); e->keep.push_back(sZ); e->keep.push_back(sr);
  if (e->h) { atx_eam_destroy(e->h); e->h = nullptr; }
  CHK(atx_eam_create_funcfl(ctx(), &sF->s, &sr->s, &sZ->s, e->cutoff, &e->h), ierror);
}
static int element_filter(const std::string &spec, const S2Particles *p, bool *ok);
static void funcfl_bind_to(void *self, void *pp, void *nn, int *ierror) {
  S2Funcfl *e = (S2Funcfl *)self;
  S2Particles *p = (S2Particles *)pp;
  S2Neighbors *n = (S2Neighbors *)nn;
  if (ierror) *ierror = ERROR_NONE;
  if (!e->h) RAISE(ierror, "The potential has not been initialised.");
  neighbors_ensure(n);
  if (!n->h || !p->h) RAISE(ierror, "No CUDA device available (no CPU fallback).");
  n->p = p;
  bool ok = true;
  const int filt = element_filter(fstr(e->elements, sizeof e->elements), p, &ok);   // filter_from_string
  if (!ok) RAISE(ierror, "Unknown element in 'elements'.");
  std::vector<int> el2db;
  for (size_t k = 0; k < p->el2Z.size(); k++) el2db.push_back((filt >> (k + 1)) & 1 ? 1 : -1);
  CHK(atx_eam_bind_to(e->h, p->h, n->h, (int)el2db.size(), el2db.data()), ierror);
}
static void funcfl_energy_and_forces(void *self, void *pp, void *nn, double *, double *epot, double *f, int *wpot_,
                                     double *mask_, double *epot_per_at, double *epot_per_bond, double *f_per_bond,
                                     double *wpot_per_at, double *wpot_per_bond, int *ierror) {
  S2Funcfl *e = (S2Funcfl *)self;
  S2Particles *p = (S2Particles *)pp;
  S2Neighbors *n = (S2Neighbors *)nn;
  if (ierror) *ierror = ERROR_NONE;
  if (epot_per_bond || f_per_bond || wpot_per_bond || mask_ || wpot_per_at)
    RAISE(ierror, "TabulatedEAM supports per-atom energies only (no masks, per-bond or per-atom virial outputs).");
  if (!e->h || !n->h) RAISE(ierror, "bind_to has not been called on this potential.");
  n->p = p;
  CHK(particles_sync(p), ierror);
  CHK(atx_eam_energy_and_forces(e->h, p->h, n->h, nullptr, epot, f, (double *)wpot_, epot_per_at, nullptr), ierror);
}

// ---- Rebo2 / Rebo2Scr (rebo2_registry.f90:22-168, rebo2_module.f90:70-223) -------------------------------
// The finished default parameter block (constants, g-splines, table coefficients) is generated at build
// time; this shim exposes `elements` and `dihedral`.  Other parameter overrides of the reference's registry
// (scalars, the nine tables) go through atomistica_b200.native.Rebo2, which rebuilds the tables.
struct S2Rebo2 {
  bool screened = false;
  char elements[1024];
  BOOL dihedral = 0;
  atx_rebo2 *h = nullptr;
  section_t *members = nullptr;
};

template <bool SCR>
static void rebo2_new(void **self, section_t *cfg, section_t **members) {
  S2Rebo2 *r = new S2Rebo2();
  r->screened = SCR;
  memset(r->elements, ' ', sizeof r->elements);
  memcpy(r->elements, "C,H", 3);
  section_t *m = ptrdict_register_section(cfg, (char *)(SCR ? "Rebo2Scr" : "Rebo2"),
                                          (char *)"The 2nd generation REBO (Brenner 2002) potential.");
  ptrdict_register_string_property(m, r->elements, (int)sizeof r->elements, (char *)"elements",
                                   (char *)"Elements for which to use this potential (default: C,H).");
  ptrdict_register_boolean_property(m, &r->dihedral, (char *)"dihedral", (char *)"Include the dihedral term?");
  r->members = m;
  *members = m;
  *self = r;
}
static void rebo2_free(void *self) {
  S2Rebo2 *r = (S2Rebo2 *)self;
  if (r->h) atx_rebo2_destroy(r->h);
  S2_CLEANUP_MEMBERS(r->members);
  delete r;
}
static void rebo2_init(void *self, int *ierror) {
  S2Rebo2 *r = (S2Rebo2 *)self;
  if (ierror) *ierror = ERROR_NONE;
  if (!ctx()) RAISE(ierror, "No CUDA device available (no CPU fallback).");
  if (fstr(r->elements, sizeof r->elements) != "C,H")
    RAISE(ierror, "Rebo2: only elements='C,H' is available through this module.");
  atx_rebo2_params par;
  atx_rebo2_screening scr;
  memset(&par, 0, sizeof par);
  memset(&scr, 0, sizeof scr);
  if (r->screened) s2_rebo2scr_defaults(&par, &scr);
  else s2_rebo2_defaults(&par, &scr);
  par.with_dihedral = r->dihedral ? 1 : 0;
  if (r->h) { atx_rebo2_destroy(r->h); r->h = nullptr; }
  if (r->screened) CHK(atx_rebo2_create_screened(ctx(), &par, &scr, &r->h), ierror);
  else CHK(atx_rebo2_create(ctx(), &par, &r->h), ierror);
}
static void rebo2_bind_to(void *self, void *pp, void *nn, int *ierror) {
  S2Rebo2 *r = (S2Rebo2 *)self;
  S2Particles *p = (S2Particles *)pp;
  S2Neighbors *n = (S2Neighbors *)nn;
  if (ierror) *ierror = ERROR_NONE;
  if (!r->h) RAISE(ierror, "The potential has not been initialised.");
  neighbors_ensure(n);
  if (!n->h || !p->h) RAISE(ierror, "No CUDA device available (no CPU fallback).");
  n->p = p;
  CHK(atx_rebo2_bind_to(r->h, p->h, n->h, (int)p->el2Z.size(), p->el2Z.data()), ierror);
}
static void rebo2_energy_and_forces(void *self, void *pp, void *nn, double *, double *epot, double *f, int *wpot_,
                                    double *mask_, double *epot_per_at, double *epot_per_bond, double *f_per_bond,
                                    double *wpot_per_at, double *wpot_per_bond, int *ierror) {
  S2Rebo2 *r = (S2Rebo2 *)self;
  S2Particles *p = (S2Particles *)pp;
  S2Neighbors *n = (S2Neighbors *)nn;
  if (ierror) *ierror = ERROR_NONE;
  if (mask_) RAISE(ierror, "Rebo2 does not support masks.");      // features: per_at, per_bond (rebo2.f90:22-27)
  if (!r->h || !n->h) RAISE(ierror, "bind_to has not been called on this potential.");
  n->p = p;
  CHK(particles_sync(p), ierror);
  CHK(atx_rebo2_energy_and_forces(r->h, p->h, n->h, epot, f, (double *)wpot_, epot_per_at, epot_per_bond, f_per_bond,
                                  wpot_per_at, wpot_per_bond), ierror);
}

// ---- pair potentials (src/potentials/pair_potentials/*.f90): el1, el2, parameters, cutoff, shift --------
struct S2Pair {
  int kind = 0;
  char el1[8], el2[8];
  double p[8];
  BOOL shift = 0;
  atx_pair *h = nullptr;
  section_t *members = nullptr;
};
struct PairSpec { int kind; const char *name; int np; const char *pn[5]; double def[5]; bool has_shift; };
static const PairSpec PAIR_SPECS[] = {
    {ATX_PAIR_LJCUT, "LJCut", 3, {"epsilon", "sigma", "cutoff"}, {0.001, 3.0, 6.0}, true},
    {ATX_PAIR_HARMONIC, "Harmonic", 3, {"k", "r0", "cutoff"}, {1.0, 1.0, 1.5}, true},
    {ATX_PAIR_DOUBLE_HARMONIC, "DoubleHarmonic", 5, {"k1", "r1", "k2", "r2", "cutoff"}, {1.0, 1.0, 1.0, 1.41421356237, 1.6}, false},
    {ATX_PAIR_BORN_MAYER, "BornMayer", 3, {"A", "rho", "cutoff"}, {1.0, 1.0, 1.0}, false},
    {ATX_PAIR_R6, "r6", 3, {"A", "r0", "cutoff"}, {1.0, 0.0, 1.0}, false},
};
template <int WHICH>
static void pair_new(void **self, section_t *cfg, section_t **members) {
  const PairSpec &sp = PAIR_SPECS[WHICH];
  S2Pair *q = new S2Pair();
  q->kind = sp.kind;
  memset(q->el1, ' ', sizeof q->el1); memset(q->el2, ' ', sizeof q->el2);
  q->el1[0] = '*'; q->el2[0] = '*';
  memset(q->p, 0, sizeof q->p);
  section_t *m = ptrdict_register_section(cfg, (char *)sp.name, (char *)"Pair potential.");
  ptrdict_register_string_property(m, q->el1, (int)sizeof q->el1, (char *)"el1", (char *)"First element.");
  ptrdict_register_string_property(m, q->el2, (int)sizeof q->el2, (char *)"el2", (char *)"Second element.");
  for (int k = 0; k < sp.np; k++) {
    q->p[k] = sp.def[k];
    ptrdict_register_real_property(m, &q->p[k], (char *)sp.pn[k], (char *)"See functional form.");
  }
  if (sp.has_shift)
    ptrdict_register_boolean_property(m, &q->shift, (char *)"shift", (char *)"Shift potential to zero energy at cutoff.");
  q->members = m;
  *members = m;
  *self = q;
}
static void pair_free(void *self) {
  S2Pair *q = (S2Pair *)self;
  if (q->h) atx_pair_destroy(q->h);
  S2_CLEANUP_MEMBERS(q->members);
  delete q;
}
static void pair_init(void *self, int *ierror) {
  S2Pair *q = (S2Pair *)self;
  if (ierror) *ierror = ERROR_NONE;
  if (!ctx()) RAISE(ierror, "No CUDA device available (no CPU fallback).");
  atx_pair_params par;
  memset(&par, 0, sizeof par);
  par.kind = q->kind;
  for (int k = 0; k < 8; k++) par.p[k] = q->p[k];
  par.shift = q->shift ? 1 : 0;
  if (q->h) { atx_pair_destroy(q->h); q->h = nullptr; }
  CHK(atx_pair_create(ctx(), &par, &q->h), ierror);
}
// filter_from_string (src/core/filter.f90:55-120): bit k = particle element id k
static int element_filter(const std::string &spec, const S2Particles *p, bool *ok) {
  *ok = true;
  int f = 0;
  if (spec == "*") {
    for (size_t k = 0; k < p->el2Z.size(); k++) f |= 1 << (k + 1);
    return f;
  }
  std::stringstream ss(spec);
  std::string sym;
  while (std::getline(ss, sym, ',')) {
    while (!sym.empty() && sym.back() == ' ') sym.pop_back();
    while (!sym.empty() && sym.front() == ' ') sym.erase(sym.begin());
    char s2[2] = {sym.size() > 0 ? sym[0] : ' ', sym.size() > 1 ? sym[1] : ' '};
    const int z = symbol_to_Z(s2);
    if (z <= 0) { *ok = false; return 0; }
    for (size_t k = 0; k < p->el2Z.size(); k++)
      if (p->el2Z[k] == z) f |= 1 << (k + 1);
  }
  return f;
}
static void pair_bind_to(void *self, void *pp, void *nn, int *ierror) {
  S2Pair *q = (S2Pair *)self;
  S2Particles *p = (S2Particles *)pp;
  S2Neighbors *n = (S2Neighbors *)nn;
  if (ierror) *ierror = ERROR_NONE;
  if (!q->h) RAISE(ierror, "The potential has not been initialised.");
  neighbors_ensure(n);
  if (!n->h || !p->h) RAISE(ierror, "No CUDA device available (no CPU fallback).");
  n->p = p;
  bool ok1, ok2;
  const int f1 = element_filter(fstr(q->el1, sizeof q->el1), p, &ok1), f2 = element_filter(fstr(q->el2, sizeof q->el2), p, &ok2);
  if (!ok1 || !ok2) RAISE(ierror, "Unknown element in el1 / el2.");
  CHK(atx_pair_bind_to(q->h, p->h, n->h, f1, f2), ierror);
}
static void pair_energy_and_forces(void *self, void *pp, void *nn, double *, double *epot, double *f, int *wpot_,
                                   double *mask_, double *epot_per_at, double *epot_per_bond, double *f_per_bond,
                                   double *wpot_per_at, double *wpot_per_bond, int *ierror) {
  S2Pair *q = (S2Pair *)self;
  S2Particles *p = (S2Particles *)pp;
  S2Neighbors *n = (S2Neighbors *)nn;
  if (ierror) *ierror = ERROR_NONE;
  if (epot_per_bond || f_per_bond || wpot_per_bond) RAISE(ierror, "This potential does not support per-bond properties.");
  if (mask_ && q->kind != ATX_PAIR_LJCUT) RAISE(ierror, "This potential does not support masks.");
  if (!q->h || !n->h) RAISE(ierror, "bind_to has not been called on this potential.");
  n->p = p;
  CHK(particles_sync(p), ierror);
  CHK(atx_pair_energy_and_forces(q->h, p->h, n->h, (int *)mask_, epot, f, (double *)wpot_, epot_per_at, wpot_per_at), ierror);
}

#ifndef ATX_SEAM2_LAMMPS   // atx_lammps.cpp includes this file and supplies the LAMMPS-flavour table instead
// the class table of src/python/c/factory.template.h; the header is generated by build.py from that
// template with N_POTENTIAL_CLASSES = the number of entries below
extern "C" {
#include "potentials_factory_c.h"
potential_class_t potential_classes[N_POTENTIAL_CLASSES] = {
#define BOP_CLASS(NAME, KIND, SCR)                                                                         \
  {NAME, bop_new<KIND, SCR>, bop_free, bop_register_data, bop_init, bop_bind_to, nullptr, nullptr, nullptr, \
   bop_energy_and_forces}
    BOP_CLASS("Tersoff", ATX_BOP_TERSOFF, false),   BOP_CLASS("TersoffScr", ATX_BOP_TERSOFF, true),
    BOP_CLASS("Kumagai", ATX_BOP_KUMAGAI, false),   BOP_CLASS("KumagaiScr", ATX_BOP_KUMAGAI, true),
    BOP_CLASS("Brenner", ATX_BOP_BRENNER, false),   BOP_CLASS("BrennerScr", ATX_BOP_BRENNER, true),
    {"TabulatedAlloyEAM", eam_new, eam_free, eam_register_data, eam_init, eam_bind_to, nullptr, nullptr, nullptr,
     eam_energy_and_forces},
    {"Rebo2", rebo2_new<false>, rebo2_free, eam_register_data, rebo2_init, rebo2_bind_to, nullptr, nullptr, nullptr,
     rebo2_energy_and_forces},
    {"Rebo2Scr", rebo2_new<true>, rebo2_free, eam_register_data, rebo2_init, rebo2_bind_to, nullptr, nullptr, nullptr,
     rebo2_energy_and_forces},
    {"Juslin", juslin_new<false>, juslin_free, eam_register_data, juslin_init, juslin_bind_to, nullptr, nullptr, nullptr,
     juslin_energy_and_forces},
    {"JuslinScr", juslin_new<true>, juslin_free, eam_register_data, juslin_init, juslin_bind_to, nullptr, nullptr,
     nullptr, juslin_energy_and_forces},
    {"TabulatedEAM", funcfl_new, funcfl_free, eam_register_data, funcfl_init, funcfl_bind_to, nullptr, nullptr, nullptr,
     funcfl_energy_and_forces},
#define PAIR_CLASS(W) \
  {"", pair_new<W>, pair_free, eam_register_data, pair_init, pair_bind_to, nullptr, nullptr, nullptr, pair_energy_and_forces}
    PAIR_CLASS(0), PAIR_CLASS(1), PAIR_CLASS(2), PAIR_CLASS(3), PAIR_CLASS(4),
};
// the pair classes take their names from PAIR_SPECS
static struct PairNames {
  PairNames() {
    for (int w = 0; w < 5; w++) strncpy(potential_classes[12 + w].name, PAIR_SPECS[w].name, MAX_NAME);
  }
} g_pair_names;
#include "coulomb_factory_c.h"
coulomb_class_t coulomb_classes[N_COULOMB_CLASSES];
}
#endif  // ATX_SEAM2_LAMMPS
