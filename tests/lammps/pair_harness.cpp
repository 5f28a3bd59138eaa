// Test harness: the call sequence of the reference's LAMMPS pair style without LAMMPS.
// Follows src/lammps/pair_style/pair_atomistica.cpp step by step -- constructor (:97-125), coeff
// (:222-276: particles_set_element per type), init_style (:281-360: class lookup, new_instance,
// optional ptrdict_read, particles_set_pointers, init, bind_to, dump_cutoffs, get_border), init_one
// (:366-384), Atomistica_neigh (:391-460: seed/last as offsets into a neighbour array based at address 0)
// and FAtomistica (:505-560: set_pointers, energy_and_forces into the live f, eng_vdwl, virial[6]).
// Linked against libatomistica_lammps.so exactly as pair_atomistica.cpp is linked against libatomistica.a.
// Test infrastructure only.
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "potentials_factory_c.h"   // the LAMMPS-flavour class table (generated from the reference's template)

extern "C" {
void particles_new(void **);
void particles_free(void *);
void particles_init(void *);
void particles_set_element(void *, const char *, int, int, int *, int *);
void particles_set_pointers(void *, int, int, int, void *, void *, void *);
void particles_get_interaction_range(void *, int, int, double *);
void particles_get_border(void *, double *);
void neighbors_new(void **);
void neighbors_free(void *);
void neighbors_init(void *);
void neighbors_set_pointers(void *, int, void *, void *, int, void *);
void neighbors_get_cutoff(void *, int, int, double *);
void neighbors_dump_cutoffs(void *, void *);
void get_full_error_string(char *);
void atomistica_startup(int);
void ptrdict_read(section_t *, char *);
void ptrdict_cleanup(section_t *);

// returns 0 or -1 (message in errbuf); out[0] = eng_vdwl, out[1..6] = virial, out[7] = rcghost, out[8] = rc
int lmp_harness_run(const char *name, const char *param_file, int ntypes, const char *const *type_symbols, int nall,
                    int nlocal, int *tag, int *type, double *x, int inum, const int *ilist, const int *numneigh,
                    int **firstneigh, int eflag_atom, int vflag_atom, int ncalls, double *f, double *eatom,
                    double *vatom, double *out, char *errbuf) {
  void *particles = nullptr, *neighbors = nullptr, *potential = nullptr;
  section_t *members = nullptr;
  int ierror = 0;
  errbuf[0] = 0;
  particles_new(&particles);
  particles_init(particles);
  neighbors_new(&neighbors);
  neighbors_init(neighbors);
  atomistica_startup(-1);
  // coeff: map LAMMPS types to elements
  for (int i = 0; i < ntypes; i++) {
    int Z;
    particles_set_element(particles, type_symbols[i], ntypes, i + 1, &Z, &ierror);
    if (ierror) { get_full_error_string(errbuf); return -1; }
  }
  // init_style
  potential_class_t *cls = nullptr;
  for (int i = 0; i < N_POTENTIAL_CLASSES; i++)
    if (!strcmp(name, potential_classes[i].name)) cls = &potential_classes[i];
  if (!cls) { snprintf(errbuf, 1000, "Could not find potential '%s' in the Atomistica potential database", name); return -1; }
  cls->new_instance(&potential, nullptr, &members);
  if (param_file && param_file[0]) ptrdict_read(members, (char *)param_file);
  particles_set_pointers(particles, nall, nlocal, nall, tag, type, x);
  cls->init(potential);
  cls->bind_to(potential, particles, neighbors, &ierror);
  if (ierror) { get_full_error_string(errbuf); return -1; }
  neighbors_dump_cutoffs(neighbors, particles);
  double rcghost = 0.0, rc = 0.0, range = 0.0;
  particles_get_border(particles, &rcghost);
  neighbors_get_cutoff(neighbors, 1, 1, &rc);
  particles_get_interaction_range(particles, 1, 1, &range);
  out[7] = rcghost;
  out[8] = rc;
  // init_one(i, j) (pair_atomistica.cpp:366-384): the per-type-pair cutoffs LAMMPS builds its lists with
  neighbors_get_cutoff(neighbors, ntypes, ntypes, &out[9]);
  neighbors_get_cutoff(neighbors, 1, ntypes, &out[10]);
  // Atomistica_neigh: seed / last relative to a neighbour array that starts at address 0
  std::vector<intptr_t> seed(nall, -1), last(nall, -2);
  int *neighb = nullptr;
  int nneighb = 0;
  for (int ii = 0; ii < inum; ii++) {
    const int i = ilist[ii];
    seed[i] = firstneigh[i] - neighb + 1;
    last[i] = seed[i] + numneigh[i] - 1;
    nneighb += numneigh[i];
  }
  double eng_vdwl = 0.0, virial[6] = {0, 0, 0, 0, 0, 0};
  for (int call = 0; call < ncalls; call++) {
    // FAtomistica
    double epot = 0.0, wpot[3][3];
    memset(wpot, 0, sizeof wpot);
    particles_set_pointers(particles, nall, nlocal, nall, tag, type, x);
    neighbors_set_pointers(neighbors, nall, seed.data(), last.data(), nneighb, neighb);
    cls->energy_and_forces(potential, particles, neighbors, &epot, f, &wpot[0][0], nullptr, eflag_atom ? eatom : nullptr,
                           vflag_atom ? vatom : nullptr, &ierror);
    if (ierror) { get_full_error_string(errbuf); return -1; }
    eng_vdwl += epot;
    virial[0] -= wpot[0][0];
    virial[1] -= wpot[1][1];
    virial[2] -= wpot[2][2];
    virial[3] -= 0.5 * (wpot[1][0] + wpot[0][1]);
    virial[4] -= 0.5 * (wpot[2][0] + wpot[0][2]);
    virial[5] -= 0.5 * (wpot[2][1] + wpot[1][2]);
  }
  out[0] = eng_vdwl;
  for (int k = 0; k < 6; k++) out[1 + k] = virial[k];
  if (members) ptrdict_cleanup(members);   // ~PairAtomistica (:137): the pair style owns the section
  cls->del(potential);
  cls->free_instance(potential);
  neighbors_free(neighbors);
  particles_free(particles);
  return 0;
}
}
