// Unscreened REBO2 per-atom device functions (the bodies of k_rebo2_bonds / k_rebo2_force of
// atx_rebo2.cu): written against plain pointers and the helpers listed in atx_rebo2_scr.cuh, so
// that tests/emu/ can compile the same source for the host and run it against the CPU oracle.
#pragma once

#include "atx_rebo2_func.cuh"

// One entry of the per-atom bond table (unscreened REBO2): 64 bytes, one cache line pair per bond
// instead of six separate arrays (the force pass walks the tables of i, j and of their neighbours;
// ncu, round 2: long-scoreboard stalls dominate k_rebo2_force_bond, L1 hit rate 37 %)
struct __align__(16) RbBond {
  double4 vec;   // unit vector i -> j and bond length
  double2 cut;   // cutoff function and its derivative
  int nb;        // neighbour (sorted numbering)
  int typ;       // pair type in bits 0..7, REBO2 element type of the neighbour (RB_C / RB_H) in bits 8..15
  int shift;     // packed periodic shift
  int slot;      // position in the atom's range of the pair list
};

__device__ __forceinline__ void rb_add3(double *f, int at, double x, double y, double z) {
  RBS_ADD(&f[3 * (size_t)at], x);
  RBS_ADD(&f[3 * (size_t)at + 1], y);
  RBS_ADD(&f[3 * (size_t)at + 2], z);
}

// w(a,b) += s * u(a) * v(b), column-major
__device__ __forceinline__ void rb_outer(double *w, double s, double ux, double uy, double uz, double vx,
                                         double vy, double vz) {
  w[0] += s * ux * vx; w[1] += s * uy * vx; w[2] += s * uz * vx;
  w[3] += s * ux * vy; w[4] += s * uy * vy; w[5] += s * uz * vy;
  w[6] += s * ux * vz; w[7] += s * uy * vz; w[8] += s * uz * vz;
}

// loop 1 + nn of one atom (bop_kernel_rebo2.f90:700-1200, SCREENING undefined)
__device__ __forceinline__ void rb_bonds_atom(int nbs, const Mat3 &A, const Rebo2Dev &P,
                                              const double4 *__restrict__ pos4,
                                              const long long *__restrict__ seed,
                                              const int2 *__restrict__ list, int *__restrict__ b_cnt,
                                              RbBond *__restrict__ b_tab,
                                              double2 *__restrict__ nn, int *__restrict__ flag, int s) {
  double4 pi = pos4[s];
  int ti = P.el2typ[(int)pi.w];
  int nb = 0;
  double nC = 0.0, nH = 0.0;
  if (ti > 0) {
    long long b0 = seed[s], b1 = seed[s + 1];
    for (long long a = b0; a < b1; a++) {
      int2 en = list[a];
      double4 pj = pos4[en.x];
      int tj = P.el2typ[(int)pj.w];
      if (tj <= 0) continue;
      double dx = pj.x - pi.x, dy = pj.y - pi.y, dz = pj.z - pi.z;
      if (ATX_NONZERO_SHIFT(en.y)) {
        int sx, sy, sz;
        atx_unpack_shift(en.y, sx, sy, sz);
        double ax, ay, az;
        atx_image_vector(A, sx, sy, sz, ax, ay, az);
        dx -= ax; dy -= ay; dz -= az;
      }
      double r2 = dx * dx + dy * dy + dz * dz;
      int ijpot = rb_Z2pair(ti, tj);
      double l = P.cut_l[ijpot];
      double fc, dfc, rl;
      if (r2 < l * l) {
        fc = 1.0; dfc = 0.0; rl = sqrt(r2);
      } else if (r2 < P.cut_h2[ijpot]) {
        rl = sqrt(r2);
        double h = P.cut_h[ijpot];
        // fCin with trig_off (rebo2_func.f90:63-85, cutoff.f90:152-196)
        if (rl > h) { fc = 0.0; dfc = 0.0; }
        else if (rl < l) { fc = 1.0; dfc = 0.0; }
        else if (rl <= l) { fc = 1.0; dfc = 0.0; }
        else if (rl >= h) { fc = 0.0; dfc = 0.0; }
        else {
          double sn, cs;
          sincos(P.cut_fac[ijpot] * (rl - l), &sn, &cs);
          fc = 0.5 * (1.0 + cs);
          dfc = -0.5 * P.cut_fac[ijpot] * sn;
        }
      } else
        continue;
      if (nb >= nbs || nb >= RB_NBL) { RBS_OR(flag, 1); break; }
      size_t q = (size_t)s * nbs + nb;
      b_tab[q].nb = en.x;
      b_tab[q].typ = ijpot | (tj << 8);
      b_tab[q].shift = en.y;
      b_tab[q].slot = (int)(a - b0);
      b_tab[q].vec = make_double4(dx / rl, dy / rl, dz / rl, rl);
      b_tab[q].cut = make_double2(fc, dfc);
      if (tj == RB_C) nC += fc; else nH += fc;
      nb++;
    }
  }
  b_cnt[s] = nb;
  nn[s] = make_double2(nC, nH);
}

// loop 2 of one atom (bop_kernel_rebo2.f90:1209-2781, SCREENING undefined, DIHEDRAL); i >= nat: nothing.
// ROLES (caller-supplied lists with explicit ghosts, role 2 = owned, 1 = ghost): every bond among the
// local atoms is evaluated and scattered as usual, but the energy and the virial of a bond are
// counted half for each OWNED end only, so that the sums over the ranks (or over the owned atoms of
// an unfolded periodic system) are the reference's totals; the caller discards the ghost rows.
// ONE: evaluate only the bond in slot ij0 of atom i (one thread per bond, k_rebo2_force_bond) instead
// of all bonds the atom is responsible for.
// NBL: depth of the per-thread scratch arrays (>= nbs); the shallow instantiation keeps the local-memory
// footprint of a resident block inside L1 for four-fold coordinated carbon
template <bool ROLES = false, bool ONE = false, int NBL = RB_NBL>
__device__ __forceinline__ void rb_force_atom(int nat, int nbs, const Rebo2Dev &P,
                                              const long long *__restrict__ seed,
                                              const int *__restrict__ b_cnt, const RbBond *__restrict__ b_tab,
                                              const double2 *__restrict__ nn,
                                              const double4 *__restrict__ pos4,
                                              const int *__restrict__ order, double *__restrict__ f,
                                              double *__restrict__ epa, double *__restrict__ wpa,
                                              double *__restrict__ epb, double *__restrict__ fpb,
                                              double *__restrict__ wpb, int i, double *acc,
                                              const unsigned char *__restrict__ role = nullptr,
                                              int ij0 = 0) {
  const int ktypi = i < nat ? P.el2typ[(int)pos4[i].w] : 0;
  const int nbi = (i < nat && ktypi > 0) ? b_cnt[i] : 0;
  if (nbi > 0) {
    const size_t qi = (size_t)i * nbs;
    double fix = 0.0, fiy = 0.0, fiz = 0.0;
    // ---- ik_loop1 (:1231-1317): conjugation inputs of the neighbours of i
    double fxik[NBL], dncx[NBL];  // fconj(x_ik), fcik * dfconj/dx
    double nconjit = 0.0;
    for (int ik = 0; ik < nbi; ik++) {
      int k = b_tab[qi + ik].nb;
      int tk = b_tab[qi + ik].typ >> 8;
      fxik[ik] = 0.0;
      dncx[ik] = 0.0;
      if (tk == RB_C) {
        double2 ck = b_tab[qi + ik].cut;
        double2 nk = nn[k];
        double xik = nk.x + nk.y - ck.x, dfx;
        rb_fconj(xik, fxik[ik], dfx);
        dncx[ik] = ck.x * dfx;
        nconjit += ck.x * fxik[ik];
      }
    }
    const double2 nni = nn[i];

    for (int ij = ONE ? ij0 : 0; ij < (ONE ? ij0 + 1 : nbi); ij++) {
      const int j = b_tab[qi + ij].nb;
      int jsx, jsy, jsz;
      atx_unpack_shift(b_tab[qi + ij].shift, jsx, jsy, jsz);
      // j_gt_i (:1332): lexicographic sign of the shift, then index
      const bool zero = (jsx == 0 && jsy == 0 && jsz == 0);
      const bool pos = jsx != 0 ? jsx > 0 : (jsy != 0 ? jsy > 0 : jsz > 0);
      // the index comparison is made in ORIGINAL atom numbering so that per-bond outputs land in
      // the same list slot as in the reference
      if (!((zero && order[j] > order[i]) || pos)) continue;
      const int ijpot = b_tab[qi + ij].typ & 255;
      const double4 vij = b_tab[qi + ij].vec;
      const double rlij = vij.w;
      if (!(rlij < P.cut_h[ijpot])) continue;
      const int ktypj = b_tab[qi + ij].typ >> 8;
      const double rlijr = 1.0 / rlij;
      const double nx = vij.x, ny = vij.y, nz = vij.z;
      const double rijx = rlij * nx, rijy = rlij * ny, rijz = rlij * nz;
      const double2 cij = b_tab[qi + ij].cut;
      const double fcarij = cij.x, dfcarijr = cij.y;
      const double2 nnj = nn[j];
      double niC = nni.x, niH = nni.y, njC = nnj.x, njH = nnj.y;
      if (ktypj == RB_C) niC -= fcarij; else niH -= fcarij;
      if (ktypi == RB_C) njC -= fcarij; else njH -= fcarij;
      if (niC > 4.0) niC = 4.0;
      if (niH > 4.0) niH = 4.0;
      double nti = niC + niH;
      if (njC > 4.0) njC = 4.0;
      if (njH > 4.0) njH = 4.0;
      double ntj = njC + njH;
      double faij, dfaijr, frij, dfrijr;
      rb_VA(P, ijpot, rlij, faij, dfaijr);
      rb_VR(P, ijpot, rlij, frij, dfrijr);
      // virial of the bond: w = -sum_a (r_a - r_i) (x) F_a over the atoms a the bond acts on.  The bond-order
      // parts are added where the FINAL force on each neighbour is known (one outer product per neighbour
      // instead of two per triplet, and no separate accumulators waiting for dfbij / dfbji)
      double wij[9];
#pragma unroll
      for (int q = 0; q < 9; q++) wij[q] = 0.0;
      double fjx = 0.0, fjy = 0.0, fjz = 0.0;
      double zij = 0.0, dix = 0, diy = 0, diz = 0, djx = 0, djy = 0, djz = 0, dzdni = 0.0;
      double nconji = 0.0;
      double dbk[NBL][3];

      // ---- ik_loop2 (:1407-1587)
      for (int ik = 0; ik < nbi; ik++) {
        const double2 cik = b_tab[qi + ik].cut;
        if (ik == ij) {
          nconji = nconjit - cik.x * fxik[ik];
          continue;
        }
        const int ikpot = b_tab[qi + ik].typ & 255;
        const double4 vik = b_tab[qi + ik].vec;
        const double rlik = vik.w;
        if (!(rlik < P.cut_h[ikpot])) {
          dbk[ik][0] = dbk[ik][1] = dbk[ik][2] = 0.0;
          continue;
        }
        const double kx = vik.x, ky = vik.y, kz = vik.z;
        const double fcik = cik.x, dfcikr = cik.y;
        double qfacan, qfadan, gfacan, gddan, dgdn;
        rb_h(P, ijpot, ikpot, rlij - rlik, qfacan, qfadan);
        const double costh = kx * nx + ky * ny + kz * nz;
        rb_g(P, ktypi, costh, nti, gfacan, gddan, dgdn);
        double ex = kx * rlik - nx * rlij, ey = ky * rlik - ny * rlij, ez = kz * rlik - nz * rlij;
        const double disjk = sqrt(ex * ex + ey * ey + ez * ez);
        ex /= disjk; ey /= disjk; ez /= disjk;
        const double dcsdij = 1.0 / rlik - costh * rlijr;
        const double dcsdik = rlijr - costh / rlik;
        const double dcsdjk = -disjk * rlijr / rlik;
        dzdni += fcik * dgdn * qfacan;
        const double dzfac = fcik * gddan * qfacan;
        zij += fcik * gfacan * qfacan;
        const double dzdrij = gfacan * fcik * qfadan;
        const double dzdrik = gfacan * (dfcikr * qfacan - fcik * qfadan);
        const double dfx = dzdrij * nx + dzfac * (dcsdij * nx - dcsdjk * ex);
        const double dfy = dzdrij * ny + dzfac * (dcsdij * ny - dcsdjk * ey);
        const double dfz = dzdrij * nz + dzfac * (dcsdij * nz - dcsdjk * ez);
        dix += -dzdrij * nx - dzdrik * kx + dzfac * (-dcsdij * nx - dcsdik * kx);
        diy += -dzdrij * ny - dzdrik * ky + dzfac * (-dcsdij * ny - dcsdik * ky);
        diz += -dzdrij * nz - dzdrik * kz + dzfac * (-dcsdij * nz - dcsdik * kz);
        djx += dfx; djy += dfy; djz += dfz;
        const double kx_ = dzdrik * kx + dzfac * (dcsdik * kx + dcsdjk * ex);
        const double ky_ = dzdrik * ky + dzfac * (dcsdik * ky + dcsdjk * ey);
        const double kz_ = dzdrik * kz + dzfac * (dcsdik * kz + dcsdjk * ez);
        dbk[ik][0] = kx_; dbk[ik][1] = ky_; dbk[ik][2] = kz_;
      }

      double pij = 0.0, dpdnci = 0.0, dpdnhi = 0.0;
      if (ktypi == RB_C) {
        rb_table2d(ijpot == RB_CC ? P.Pcc : P.Pch, 5, 5, niH, niC, pij, dpdnhi, dpdnci);
        zij += pij;
        dpdnci += dzdni;
        dpdnhi += dzdni;
      }
      double bij, dfbij;
      rb_bo(P, ktypi, zij, fcarij, faij, bij, dfbij);

      // ---- jl_loop (:1644-1887)
      const size_t qj = (size_t)j * nbs;
      const int nbj = b_cnt[j];
      double zji = 0.0, bix = 0, biy = 0, biz = 0, bjx = 0, bjy = 0, bjz = 0, dzdnj = 0.0;
      double nconjj = 0.0;
      double dbl[NBL][3];
      double fxjl[NBL], dnlx[NBL];
      for (int jl = 0; jl < nbj; jl++) {
        fxjl[jl] = 0.0; dnlx[jl] = 0.0;
        dbl[jl][0] = dbl[jl][1] = dbl[jl][2] = 0.0;
        const int l = b_tab[qj + jl].nb;
        int lsx, lsy, lsz;
        atx_unpack_shift(b_tab[qj + jl].shift, lsx, lsy, lsz);
        lsx += jsx; lsy += jsy; lsz += jsz;
        if (l == i && lsx == 0 && lsy == 0 && lsz == 0) continue;   // l_neq_i
        const int ktypl = b_tab[qj + jl].typ >> 8;
        const int jlpot = b_tab[qj + jl].typ & 255;
        const double4 vjl = b_tab[qj + jl].vec;
        const double rljl = vjl.w;
        const double lx = vjl.x, ly = vjl.y, lz = vjl.z;
        const double2 cjl = b_tab[qj + jl].cut;
        const double fcjl = cjl.x, dfcjlr = cjl.y;
        if (ktypl == RB_C) {
          double2 nl_ = nn[l];
          double xjl = nl_.x + nl_.y - fcjl, dfx;
          rb_fconj(xjl, fxjl[jl], dfx);
          dnlx[jl] = fcjl * dfx;
          nconjj += fcjl * fxjl[jl];
        }
        if (rljl < P.cut_h[jlpot]) {
          double qfacan, qfadan, gfacan, gddan, dgdn;
          rb_h(P, ijpot, jlpot, rlij - rljl, qfacan, qfadan);
          const double costh = -(lx * nx + ly * ny + lz * nz);
          rb_g(P, ktypj, costh, ntj, gfacan, gddan, dgdn);
          double ex = lx * rljl + nx * rlij, ey = ly * rljl + ny * rlij, ez = lz * rljl + nz * rlij;
          const double disil = sqrt(ex * ex + ey * ey + ez * ez);
          ex /= disil; ey /= disil; ez /= disil;
          const double dcsdji = 1.0 / rljl - costh * rlijr;
          const double dcsdjl = rlijr - costh / rljl;
          const double dcsdil = -disil * rlijr / rljl;
          dzdnj += fcjl * dgdn * qfacan;
          const double dzfac = fcjl * gddan * qfacan;
          zji += fcjl * gfacan * qfacan;
          const double dzdrji = gfacan * fcjl * qfadan;
          const double dzdrjl = gfacan * (dfcjlr * qfacan - fcjl * qfadan);
          bjx += dzdrji * nx - dzdrjl * lx + dzfac * (dcsdji * nx - dcsdjl * lx);
          bjy += dzdrji * ny - dzdrjl * ly + dzfac * (dcsdji * ny - dcsdjl * ly);
          bjz += dzdrji * nz - dzdrjl * lz + dzfac * (dcsdji * nz - dcsdjl * lz);
          const double dfx = -dzdrji * nx + dzfac * (-dcsdji * nx - dcsdil * ex);
          const double dfy = -dzdrji * ny + dzfac * (-dcsdji * ny - dcsdil * ey);
          const double dfz = -dzdrji * nz + dzfac * (-dcsdji * nz - dcsdil * ez);
          bix += dfx; biy += dfy; biz += dfz;
          const double lx_ = dzdrjl * lx + dzfac * (dcsdjl * lx + dcsdil * ex);
          const double ly_ = dzdrjl * ly + dzfac * (dcsdjl * ly + dcsdil * ey);
          const double lz_ = dzdrjl * lz + dzfac * (dcsdjl * lz + dcsdil * ez);
          dbl[jl][0] = lx_; dbl[jl][1] = ly_; dbl[jl][2] = lz_;
        }
      }

      double pji = 0.0, dpdncj = 0.0, dpdnhj = 0.0;
      if (ktypj == RB_C) {
        rb_table2d(ijpot == RB_CC ? P.Pcc : P.Pch, 5, 5, njH, njC, pji, dpdnhj, dpdncj);
        zji += pji;
        dpdncj += dzdnj;
        dpdnhj += dzdnj;
      }
      double bji, dfbji;
      rb_bo(P, ktypj, zji, fcarij, faij, bji, dfbji);

      double nconj = nconji * nconji + nconjj * nconjj;
      if (nconj > 8.0) nconj = 8.0;
      if (nti > 3.0) nti = 3.0;
      if (ntj > 3.0) ntj = 3.0;

      // ---- dihedral (:1950-2087), only when with_dihedral
      double bdh = 0.0, tij = 0.0, dtdni = 0.0, dtdnj = 0.0, dtdncn = 0.0;
      if (P.with_dihedral && ijpot == RB_CC) {
        rb_table3d(P.Tcc, 4, 4, 9, nti, ntj, nconj, tij, dtdni, dtdnj, dtdncn);
        const double tije = tij * faij * fcarij;
        if (tij != 0) {
          for (int ik = 0; ik < nbi; ik++) {
            if (ik == ij) continue;
            const int k = b_tab[qi + ik].nb;
            int ksx, ksy, ksz;
            atx_unpack_shift(b_tab[qi + ik].shift, ksx, ksy, ksz);
            const double4 vik = b_tab[qi + ik].vec;
            const double rlik = vik.w, kx = vik.x, ky = vik.y, kz = vik.z;
            const double2 cik = b_tab[qi + ik].cut;
            const double fcik = cik.x, dfcikr = cik.y;
            const double dot_ij_ik = nx * kx + ny * ky + nz * kz;
            const double dcik = 1.0 - dot_ij_ik * dot_ij_ik;
            for (int jl = 0; jl < nbj; jl++) {
              const int l = b_tab[qj + jl].nb;
              int lsx, lsy, lsz;
              atx_unpack_shift(b_tab[qj + jl].shift, lsx, lsy, lsz);
              lsx += jsx; lsy += jsy; lsz += jsz;
              if (l == i && lsx == 0 && lsy == 0 && lsz == 0) continue;
              if (l == k && lsx == ksx && lsy == ksy && lsz == ksz) continue;
              const double4 vjl = b_tab[qj + jl].vec;
              const double rljl = vjl.w, lx = vjl.x, ly = vjl.y, lz = vjl.z;
              const double2 cjl = b_tab[qj + jl].cut;
              const double fcjl = cjl.x, dfcjlr = cjl.y;
              const double dot_ij_jl = nx * lx + ny * ly + nz * lz;
              const double dot_ik_jl = kx * lx + ky * ly + kz * lz;
              const double dcjl = 1.0 - dot_ij_jl * dot_ij_jl;
              const double abs_dc = sqrt(dcik * dcjl);
              const double cost = (dot_ij_ik * dot_ij_jl - dot_ik_jl) / abs_dc;
              double bdhij = 1 - cost * cost;
              bdh += bdhij * fcik * fcjl;
              bdhij = bdhij * tij * faij * fcarij / 2;
              const double dbdhij = -2 * cost * tije * fcik * fcjl / 2;
              const double a1 = dot_ij_jl / abs_dc + cost * dot_ij_ik / dcik;
              const double a2 = dot_ij_ik / abs_dc + cost * dot_ij_jl / dcjl;
              const double a3 = 2 * dot_ik_jl / abs_dc + cost * (1.0 / dcik + 1.0 / dcjl);
              double dx_ = dbdhij * (a1 * kx + a2 * lx - a3 * nx) / rlij;
              double dy_ = dbdhij * (a1 * ky + a2 * ly - a3 * ny) / rlij;
              double dz_ = dbdhij * (a1 * kz + a2 * lz - a3 * nz) / rlij;
              fix += dx_; fiy += dy_; fiz += dz_;
              fjx -= dx_; fjy -= dy_; fjz -= dz_;
              rb_outer(wij, 1.0, rijx, rijy, rijz, dx_, dy_, dz_);
              dx_ = dbdhij * (-1.0 / dcik * cost * kx - 1.0 / abs_dc * lx + a1 * nx) / rlik + bdhij * dfcikr * fcjl * kx;
              dy_ = dbdhij * (-1.0 / dcik * cost * ky - 1.0 / abs_dc * ly + a1 * ny) / rlik + bdhij * dfcikr * fcjl * ky;
              dz_ = dbdhij * (-1.0 / dcik * cost * kz - 1.0 / abs_dc * lz + a1 * nz) / rlik + bdhij * dfcikr * fcjl * kz;
              fix += dx_; fiy += dy_; fiz += dz_;
              rb_add3(f, k, -dx_, -dy_, -dz_);
              rb_outer(wij, 1.0, rlik * kx, rlik * ky, rlik * kz, dx_, dy_, dz_);
              dx_ = dbdhij * (-1.0 / dcjl * cost * lx - 1.0 / abs_dc * kx + a2 * nx) / rljl + bdhij * fcik * dfcjlr * lx;
              dy_ = dbdhij * (-1.0 / dcjl * cost * ly - 1.0 / abs_dc * ky + a2 * ny) / rljl + bdhij * fcik * dfcjlr * ly;
              dz_ = dbdhij * (-1.0 / dcjl * cost * lz - 1.0 / abs_dc * kz + a2 * nz) / rljl + bdhij * fcik * dfcjlr * lz;
              fjx += dx_; fjy += dy_; fjz += dz_;
              rb_add3(f, l, -dx_, -dy_, -dz_);
              rb_outer(wij, 1.0, rljl * lx, rljl * ly, rljl * lz, dx_, dy_, dz_);
            }
          }
        }
      }

      double fij = 0.0, dfdni = 0.0, dfdnj = 0.0, dfdncn = 0.0;
      if (ijpot == RB_CC) rb_table3d(P.Fcc, 4, 4, 9, nti, ntj, nconj, fij, dfdni, dfdnj, dfdncn);
      else if (ijpot == RB_HH) rb_table3d(P.Fhh, 4, 4, 9, nti, ntj, nconj, fij, dfdni, dfdnj, dfdncn);
      else if (ktypi == RB_C) rb_table3d(P.Fch, 4, 4, 9, ntj, nti, nconj, fij, dfdnj, dfdni, dfdncn);
      else if (ktypj == RB_C) rb_table3d(P.Fch, 4, 4, 9, nti, ntj, nconj, fij, dfdni, dfdnj, dfdncn);
      dfdni += dtdni * bdh;
      dfdnj += dtdnj * bdh;
      dfdncn += dtdncn * bdh;
      dfdni = 0.5 * fcarij * faij * dfdni;
      dfdnj = 0.5 * fcarij * faij * dfdnj;
      dfdncn = 0.5 * fcarij * faij * dfdncn;
      const double dfdncni = 2 * dfdncn * nconji;
      const double dfdncnj = 2 * dfdncn * nconjj;

      // ---- forces through N_i, N^conj_i on the neighbours k of i and their neighbours m (:2433-2470)
      for (int ik = 0; ik < nbi; ik++) {
        if (ik == ij) continue;
        const int k = b_tab[qi + ik].nb;
        const int tk = b_tab[qi + ik].typ >> 8;
        const double4 vik = b_tab[qi + ik].vec;
        const double2 cik = b_tab[qi + ik].cut;
        // dnidk(:, ikc, type) = rnik*dfcikr for the type of k, 0 for the other type
        const double sC = (tk == RB_C) ? cik.y : 0.0, sH = (tk == RB_H) ? cik.y : 0.0;
        const double dncdk = fxik[ik] * cik.y;  // dncnidk = nconjdr * rnik (0 unless k is C)
        const double pref = -(dfdni * (sC + sH) + dfdncni * dncdk) - dfbij * (dpdnci * sC + dpdnhi * sH);
        const double dx_ = pref * vik.x, dy_ = pref * vik.y, dz_ = pref * vik.z;
        double fkx = dx_, fky = dy_, fkz = dz_;
        fix -= dx_; fiy -= dy_; fiz -= dz_;
        rb_outer(wij, -1.0, vik.w * vik.x, vik.w * vik.y, vik.w * vik.z, dx_, dy_, dz_);
        if (tk == RB_C && dfdncni * dncx[ik] != 0.0) {
          int ksx, ksy, ksz;
          atx_unpack_shift(b_tab[qi + ik].shift, ksx, ksy, ksz);
          const size_t qk = (size_t)k * nbs;
          const int nbk = b_cnt[k];
          for (int km = 0; km < nbk; km++) {
            const int m = b_tab[qk + km].nb;
            int msx, msy, msz;
            atx_unpack_shift(b_tab[qk + km].shift, msx, msy, msz);
            if (m == i && msx + ksx == 0 && msy + ksy == 0 && msz + ksz == 0) continue;
            const double4 vkm = b_tab[qk + km].vec;
            const double c = -dfdncni * dncx[ik] * b_tab[qk + km].cut.y;
            const double mx = c * vkm.x, my = c * vkm.y, mz = c * vkm.z;
            rb_add3(f, m, mx, my, mz);
            fkx -= mx; fky -= my; fkz -= mz;
            rb_outer(wij, -1.0, vkm.w * vkm.x, vkm.w * vkm.y, vkm.w * vkm.z, mx, my, mz);
          }
        }
        // bond-order force on k (:2650-2662) and its share of the virial
        fkx += -dfbij * dbk[ik][0]; fky += -dfbij * dbk[ik][1]; fkz += -dfbij * dbk[ik][2];
        rb_outer(wij, dfbij, vik.w * vik.x, vik.w * vik.y, vik.w * vik.z, dbk[ik][0], dbk[ik][1], dbk[ik][2]);
        rb_add3(f, k, fkx, fky, fkz);
      }
      // ---- same on the j side (:2472-2517)
      for (int jl = 0; jl < nbj; jl++) {
        const int l = b_tab[qj + jl].nb;
        int lsx, lsy, lsz;
        atx_unpack_shift(b_tab[qj + jl].shift, lsx, lsy, lsz);
        lsx += jsx; lsy += jsy; lsz += jsz;
        if (l == i && lsx == 0 && lsy == 0 && lsz == 0) continue;
        const int tl = b_tab[qj + jl].typ >> 8;
        const double4 vjl = b_tab[qj + jl].vec;
        const double2 cjl = b_tab[qj + jl].cut;
        const double sC = (tl == RB_C) ? cjl.y : 0.0, sH = (tl == RB_H) ? cjl.y : 0.0;
        const double dncdl = fxjl[jl] * cjl.y;
        const double pref = -(dfdnj * (sC + sH) + dfdncnj * dncdl) - dfbji * (dpdncj * sC + dpdnhj * sH);
        const double dx_ = pref * vjl.x, dy_ = pref * vjl.y, dz_ = pref * vjl.z;
        double flx = dx_, fly = dy_, flz = dz_;
        fjx -= dx_; fjy -= dy_; fjz -= dz_;
        rb_outer(wij, -1.0, vjl.w * vjl.x, vjl.w * vjl.y, vjl.w * vjl.z, dx_, dy_, dz_);
        if (tl == RB_C && dfdncnj * dnlx[jl] != 0.0) {
          const size_t ql = (size_t)l * nbs;
          const int nbl = b_cnt[l];
          for (int ln = 0; ln < nbl; ln++) {
            const int n = b_tab[ql + ln].nb;
            int nsx, nsy, nsz;
            atx_unpack_shift(b_tab[ql + ln].shift, nsx, nsy, nsz);
            // n /= j .or. ndc /= jdc with ndc = ldc + dcell(ln)
            if (n == j && nsx + lsx == jsx && nsy + lsy == jsy && nsz + lsz == jsz) continue;
            const double4 vln = b_tab[ql + ln].vec;
            const double c = -dfdncnj * dnlx[jl] * b_tab[ql + ln].cut.y;
            const double mx = c * vln.x, my = c * vln.y, mz = c * vln.z;
            rb_add3(f, n, mx, my, mz);
            flx -= mx; fly -= my; flz -= mz;
            rb_outer(wij, -1.0, vln.w * vln.x, vln.w * vln.y, vln.w * vln.z, mx, my, mz);
          }
        }
        flx += -dfbji * dbl[jl][0]; fly += -dfbji * dbl[jl][1]; flz += -dfbji * dbl[jl][2];
        rb_outer(wij, dfbji, vjl.w * vjl.x, vjl.w * vjl.y, vjl.w * vjl.z, dbl[jl][0], dbl[jl][1], dbl[jl][2]);
        rb_add3(f, l, flx, fly, flz);
      }

      // ---- pair terms (:2525-2716)
      const double baveij = 0.5 * (bij + bji + fij + tij * bdh);
      const double hlfvij = fcarij * (frij + baveij * faij) / 2;
      double wown = 1.0;  // share of this bond that belongs to owned atoms
      if (ROLES) {
        wown = 0.5 * ((role[i] >= 2 ? 1.0 : 0.0) + (role[j] >= 2 ? 1.0 : 0.0));
        acc[0] += 2 * hlfvij * wown;
      } else
        acc[0] += 2 * hlfvij;
      if (epa) {
        RBS_ADD(&epa[i], hlfvij);
        RBS_ADD(&epa[j], hlfvij);
      }
      const double dffac = dfrijr * fcarij + baveij * dfaijr * fcarij + frij * dfcarijr + baveij * faij * dfcarijr;
      const double dfx = dffac * nx, dfy = dffac * ny, dfz = dffac * nz;
      fix += dfx; fiy += dfy; fiz += dfz;
      fjx -= dfx; fjy -= dfy; fjz -= dfz;
      rb_outer(wij, 1.0, rijx, rijy, rijz, dfx, dfy, dfz);
      rb_outer(wij, dfbij, rijx, rijy, rijz, djx, djy, djz);
      rb_outer(wij, -dfbji, rijx, rijy, rijz, bix, biy, biz);
      fix += -(dfbij * dix + dfbji * bix); fiy += -(dfbij * diy + dfbji * biy); fiz += -(dfbij * diz + dfbji * biz);
      fjx += -(dfbij * djx + dfbji * bjx); fjy += -(dfbij * djy + dfbji * bjy); fjz += -(dfbij * djz + dfbji * bjz);
      rb_add3(f, j, fjx, fjy, fjz);
#pragma unroll
      for (int q = 0; q < 9; q++) acc[1 + q] += ROLES ? wown * wij[q] : wij[q];
      const long long a = seed[i] + b_tab[qi + ij].slot;
      if (epb) epb[a] = 2 * hlfvij;
      if (fpb) { fpb[3 * a] = dfx; fpb[3 * a + 1] = dfy; fpb[3 * a + 2] = dfz; }
      if (wpb) {
#pragma unroll
        for (int q = 0; q < 9; q++) wpb[9 * a + q] = wij[q];
      }
      if (wpa) {
#pragma unroll
        for (int q = 0; q < 9; q++) {
          RBS_ADD(&wpa[9 * (size_t)i + q], 0.5 * wij[q]);
          RBS_ADD(&wpa[9 * (size_t)j + q], 0.5 * wij[q]);
        }
      }
    }
    rb_add3(f, i, fix, fiy, fiz);
  }
}

// The bonds atom i is responsible for: exactly the slots that the ij loop of rb_force_atom does not
// skip (j_gt_i in original numbering, bop_kernel_rebo2.f90:1332, and rlij < cut_h).  Returns their
// number and writes (i, slot) pairs to `out` when it is given.
__device__ __forceinline__ int rb_owned_bonds(int nbs, const Rebo2Dev &P, const int *__restrict__ b_cnt,
                                              const RbBond *__restrict__ b_tab,
                                              const double4 *__restrict__ pos4,
                                              const int *__restrict__ order, int i, int2 *out) {
  const int ktypi = P.el2typ[(int)pos4[i].w];
  const int nbi = ktypi > 0 ? b_cnt[i] : 0;
  const size_t qi = (size_t)i * nbs;
  int n = 0;
  for (int ij = 0; ij < nbi; ij++) {
    const int j = b_tab[qi + ij].nb;
    int jsx, jsy, jsz;
    atx_unpack_shift(b_tab[qi + ij].shift, jsx, jsy, jsz);
    const bool zero = (jsx == 0 && jsy == 0 && jsz == 0);
    const bool pos = jsx != 0 ? jsx > 0 : (jsy != 0 ? jsy > 0 : jsz > 0);
    if (!((zero && order[j] > order[i]) || pos)) continue;
    if (!(b_tab[qi + ij].vec.w < P.cut_h[b_tab[qi + ij].typ & 255])) continue;
    if (out) out[n] = make_int2(i, ij);
    n++;
  }
  return n;
}
