// Minimal stand-ins for the LAMMPS classes src/lammps/pair_style/pair_atomistica.cpp of the reference uses, so
// that the UNMODIFIED pair style source compiles and runs without LAMMPS (tests/test_seam2_lammps.py).  Only
// the members that file touches exist.  Test infrastructure; written for this repository (LAMMPS itself is not
// in the image).
#ifndef ATX_LMP_STUB_H
#define ATX_LMP_STUB_H
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <stdexcept>
#include <string>

#define FLERR __FILE__, __LINE__
#ifndef MAX
#define MAX(a, b) ((a) > (b) ? (a) : (b))
#endif
#ifndef MIN
#define MIN(a, b) ((a) < (b) ? (a) : (b))
#endif

namespace LAMMPS_NS {

class Pair;

namespace NeighConst {
enum { REQ_DEFAULT = 0, REQ_FULL = 1 << 0, REQ_GHOST = 1 << 1 };
}

class Error {
 public:
  void all(const char *file, int line, const char *msg) { throw std::runtime_error(fmt(file, line, msg)); }
  void all(const char *file, int line, const std::string &msg) { all(file, line, msg.c_str()); }
  void one(const char *file, int line, const char *msg) { throw std::runtime_error(fmt(file, line, msg)); }

 private:
  static std::string fmt(const char *file, int line, const char *msg) {
    return std::string(msg) + " (" + file + ":" + std::to_string(line) + ")";
  }
};

class Memory {
 public:
  template <typename T>
  T **create(T **&array, int n1, int n2, const char *) {
    T *data = (T *)calloc((size_t)n1 * n2, sizeof(T));
    array = (T **)malloc(sizeof(T *) * n1);
    for (int i = 0; i < n1; i++) array[i] = data + (size_t)i * n2;
    return array;
  }
  template <typename T>
  void destroy(T **&array) {
    if (!array) return;
    free(array[0]);
    free(array);
    array = nullptr;
  }
  void *smalloc(size_t n, const char *) { return n ? malloc(n) : nullptr; }
  void sfree(void *p) { free(p); }
};

class Atom {
 public:
  double **x = nullptr, **f = nullptr;
  int *type = nullptr, *tag = nullptr, *gid = nullptr;
  int nlocal = 0, nghost = 0, nmax = 0, ntypes = 0, tag_enable = 1, gfmd_flag = 0;
};

class Force {
 public:
  int newton_pair = 1;
};

class Comm {
 public:
  double cutghostuser = 0.0;
};

class Neighbor {
 public:
  Pair *requestor = nullptr;
  int flags = 0;
  void add_request(Pair *p, int f) { requestor = p; flags = f; }
};

class NeighList {
 public:
  int inum = 0, gnum = 0, ghost = 1;
  int *ilist = nullptr, *numneigh = nullptr;
  int **firstneigh = nullptr;
};

class Update {
 public:
  const char *unit_style = "metal";
};

class LAMMPS {
 public:
  Memory *memory;
  Error *error;
  Atom *atom;
  Force *force;
  Comm *comm;
  Neighbor *neighbor;
  Update *update;
};

class Pointers {
 public:
  explicit Pointers(LAMMPS *ptr)
      : lmp(ptr), memory(ptr->memory), error(ptr->error), atom(ptr->atom), force(ptr->force), comm(ptr->comm),
        neighbor(ptr->neighbor), update(ptr->update) {}
  virtual ~Pointers() {}

 protected:
  LAMMPS *lmp;
  Memory *&memory;
  Error *&error;
  Atom *&atom;
  Force *&force;
  Comm *&comm;
  Neighbor *&neighbor;
  Update *&update;
};

// the part of LAMMPS' Pair the pair style relies on (pair.h / pair.cpp: ev_setup zeroes the accumulators and
// sizes the per-atom arrays; eflag / vflag bits as in LAMMPS: 1 = global, 2 = per atom (energy), 4 = per atom (virial))
class Pair : protected Pointers {
 public:
  explicit Pair(LAMMPS *lmp) : Pointers(lmp) {}
  ~Pair() override {
    free(eatom);
    if (vatom) { free(vatom[0]); free(vatom); }
  }
  virtual void compute(int, int) = 0;
  virtual void settings(int, char **) = 0;
  virtual void coeff(int, char **) = 0;
  virtual void init_style() {}
  virtual double init_one(int, int) { return 0.0; }
  virtual double memory_usage() { return 0.0; }

  int single_enable = 1, one_coeff = 0, no_virial_fdotr_compute = 0, ghostneigh = 0, allocated = 0;
  int evflag = 0, eflag_either = 0, eflag_global = 0, eflag_atom = 0, vflag_either = 0, vflag_global = 0,
      vflag_atom = 0, vflag_fdotr = 0;
  int **setflag = nullptr;
  double **cutsq = nullptr, **cutghost = nullptr;
  double eng_vdwl = 0.0, eng_coul = 0.0, virial[6] = {0, 0, 0, 0, 0, 0};
  double *eatom = nullptr, **vatom = nullptr;
  NeighList *list = nullptr;

 protected:
  int maxeatom = 0, maxvatom = 0;
  void ev_setup(int eflag, int vflag) {
    evflag = 1;
    eflag_either = eflag;
    eflag_global = eflag & 1;
    eflag_atom = eflag & 2;
    vflag_either = vflag;
    vflag_global = vflag & 3;
    vflag_atom = vflag & 4;
    const int n = atom->nmax;
    if (eflag_atom && n > maxeatom) {
      free(eatom);
      eatom = (double *)malloc(sizeof(double) * n);
      maxeatom = n;
    }
    if (vflag_atom && n > maxvatom) {
      if (vatom) { free(vatom[0]); free(vatom); }
      double *d = (double *)malloc(sizeof(double) * 6 * n);
      vatom = (double **)malloc(sizeof(double *) * n);
      for (int i = 0; i < n; i++) vatom[i] = d + 6 * (size_t)i;
      maxvatom = n;
    }
    if (eflag_global) eng_vdwl = eng_coul = 0.0;
    if (vflag_global)
      for (int k = 0; k < 6; k++) virial[k] = 0.0;
    const int nall = atom->nlocal + atom->nghost;
    if (eflag_atom)
      for (int i = 0; i < nall; i++) eatom[i] = 0.0;
    if (vflag_atom)
      for (int i = 0; i < nall; i++)
        for (int k = 0; k < 6; k++) vatom[i][k] = 0.0;
  }
};

}  // namespace LAMMPS_NS
#endif
