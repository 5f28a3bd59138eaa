// Drives the reference's UNMODIFIED LAMMPS pair style (src/lammps/pair_style/pair_atomistica.cpp, compiled from
// where it lies in the reference tree against the stand-in headers of tests/lammps/stub/) the way LAMMPS does:
// pair_style atomistica <name> [<file>] -> settings(); pair_coeff * * El1 El2 ... -> coeff(); init_style();
// init_one(i, j) for every type pair; then compute(eflag, vflag) with a full neighbour list that includes
// ghost atoms.  Same C entry point signature as lmp_harness_run (pair_harness.cpp), so the tests run both.
// Linked against libatomistica_lammps.so exactly as the pair style is linked against libatomistica.a.
// Test infrastructure only.
#include <string>
#include <vector>

#include "pair_atomistica.h"

using namespace LAMMPS_NS;

extern "C" int lmp_pair_run(const char *name, const char *param_file, int ntypes, const char *const *type_symbols,
                            int nall, int nlocal, int *tag, int *type, double *x, int inum, const int *ilist,
                            const int *numneigh, int **firstneigh, int eflag_atom, int vflag_atom, int ncalls,
                            double *f, double *eatom, double *vatom, double *out, char *errbuf) {
  errbuf[0] = 0;
  Memory memory;
  Error error;
  Atom atom;
  Force force;
  Comm comm;
  Neighbor neighbor;
  Update update;
  LAMMPS lmp{&memory, &error, &atom, &force, &comm, &neighbor, &update};
  std::vector<double *> xr(nall), fr(nall);
  for (int i = 0; i < nall; i++) { xr[i] = x + 3 * (size_t)i; fr[i] = f + 3 * (size_t)i; }
  atom.x = xr.data();
  atom.f = fr.data();
  atom.type = type;
  atom.tag = tag;
  atom.nlocal = nlocal;
  atom.nghost = nall - nlocal;
  atom.nmax = nall;
  atom.ntypes = ntypes;
  NeighList list;
  list.inum = inum < nlocal ? inum : nlocal;     // owned atoms first, then the ghost atoms that carry lists
  list.gnum = inum - list.inum;
  list.ghost = 1;
  list.ilist = const_cast<int *>(ilist);
  list.numneigh = const_cast<int *>(numneigh);
  list.firstneigh = firstneigh;
  try {
    PairAtomistica pair(&lmp);
    std::vector<std::string> sa{name};
    if (param_file && param_file[0]) sa.push_back(param_file);
    std::vector<char *> sp;
    for (auto &s : sa) sp.push_back(&s[0]);
    pair.settings((int)sp.size(), sp.data());
    std::vector<std::string> ca{"*", "*"};
    for (int i = 0; i < ntypes; i++) ca.push_back(type_symbols[i]);
    ca.push_back("");   // coeff() writes one int past its map[ntypes] for the last type (pair_atomistica.cpp:244)
    std::vector<char *> cp;
    for (auto &s : ca) cp.push_back(&s[0]);
    pair.coeff(2 + ntypes, cp.data());
    pair.init_style();
    if (!(neighbor.flags & NeighConst::REQ_FULL) || !(neighbor.flags & NeighConst::REQ_GHOST))
      throw std::runtime_error("the pair style did not request a full list with ghost atoms");
    for (int i = 1; i <= ntypes; i++)
      for (int j = i; j <= ntypes; j++) pair.init_one(i, j);
    out[7] = comm.cutghostuser;
    out[8] = pair.init_one(1, 1);
    out[9] = pair.init_one(ntypes, ntypes);
    out[10] = pair.init_one(1, ntypes);
    pair.list = &list;
    double eng = 0.0, virial[6] = {0, 0, 0, 0, 0, 0};
    for (int call = 0; call < ncalls; call++) {
      pair.compute(1 | (eflag_atom ? 2 : 0), 1 | (vflag_atom ? 4 : 0));
      eng += pair.eng_vdwl;                        // ev_setup zeroes the accumulators in every call
      for (int k = 0; k < 6; k++) virial[k] += pair.virial[k];
      if (eflag_atom)
        for (int i = 0; i < nall; i++) eatom[i] += pair.eatom[i];
      if (vflag_atom)
        for (int i = 0; i < nall; i++)
          for (int k = 0; k < 6; k++) vatom[6 * (size_t)i + k] += pair.vatom[i][k];
    }
    out[0] = eng;
    for (int k = 0; k < 6; k++) out[1 + k] = virial[k];
  } catch (const std::exception &e) {
    snprintf(errbuf, 1000, "%s", e.what());
    return -1;
  }
  return 0;
}

// The ownership sequence of ~PairAtomistica (pair_atomistica.cpp:137-142) and of init_style's re-initialisation
// (:296-306) for one class, without init (so it runs without a GPU): the pair style cleans the ptrdict section
// ITSELF, then calls del and free_instance -- an instance that cleaned the section again would free it twice.
extern "C" int lmp_pair_ownership(const char *name) {
  potential_class_t *cls = nullptr;
  for (int i = 0; i < N_POTENTIAL_CLASSES; i++)
    if (!strcmp(name, potential_classes[i].name)) cls = &potential_classes[i];
  if (!cls) return -1;
  for (int rep = 0; rep < 3; rep++) {
    void *pot = nullptr;
    section_t *members = nullptr;
    cls->new_instance(&pot, nullptr, &members);
    if (!pot || !members) return -2;
    ptrdict_cleanup(members);
    cls->del(pot);
    cls->free_instance(pot);
  }
  return 0;
}
