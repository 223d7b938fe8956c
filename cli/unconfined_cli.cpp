// unconfined_cli -- the reference's command line on top of the C ABI:
//     unconfined_cli [input.dat] [--fresh-abscissae] [--ngpu N] [--stdout] [--dump | --header-only]
// reads the deck exactly as read_input does (driver_io.f90:30-666), evaluates
// driver.f90:100-231 through ONE unc_eval_grid call on the GPU(s), and writes the
// reference's output file (headers driver_io.f90:668-845, rows driver.f90:234-272).
// There is no CPU evaluation path: without a CUDA device the call fails with the
// library's error.  --dump prints the parsed/derived deck (hex floats) and exits without
// touching the device (used by the CPU tests).
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <iostream>
#include <stdexcept>
#include <string>
#include <vector>

#include "../include/unconfined_b200.h"
#include "deck_io.hpp"

using uncli::Deck;

static void fill_tables(Deck &d) {
  // driver_io.f90:572-586, 628-664 through the library's host routines
  const int terms = std::max(d.j0s[0], d.j0s[1]) + d.nacc + 1;
  d.j0z.resize(terms);
  if (unc_j0_zeros(terms, d.j0z.data())) throw std::runtime_error(unc_last_error());
  d.sv.resize(d.tD.size());
  if (unc_split_index((int32_t)d.tD.size(), d.tD.data(), d.j0s[0], d.j0s[1], d.sv.data()))
    throw std::runtime_error(unc_last_error());
  d.zLay.resize(d.zD.size());
  if (unc_zlay((int32_t)d.zD.size(), d.zD.data(), d.lD, d.dD, d.zLay.data())) throw std::runtime_error(unc_last_error());
}

static unc_params make_params(const Deck &d) {
  unc_params p;
  std::memset(&p, 0, sizeof p);
  p.model = d.model; p.M = d.M; p.alpha = d.alpha; p.tol = d.tol;
  p.tee_mult = 2.0;                                   // driver.f90:54
  p.time_type = d.timeType; p.n_time_par = (int32_t)d.timePar.size(); p.time_par = d.timePar.data();
  p.ts_k = d.ts_k; p.ts_R = d.ts_R; p.gl_nacc = d.nacc; p.gl_ord = d.ord;
  p.n_j0z = (int32_t)d.j0z.size(); p.j0z = d.j0z.data();
  p.moench_M = d.MoenchM; p.moench_gamma = d.MoenchM ? d.MoenchGamma.data() : nullptr;
  p.kappa = d.kappa; p.alphaD = d.alphaD; p.beta = d.beta;
  p.lD = d.lD; p.dD = d.dD; p.bD = d.bD; p.rDw = d.rDw;
  p.l = d.l; p.d = d.d; p.Ss = d.Ss; p.rDwobs = d.rDwobs; p.sF = d.sF;
  p.mn_type = d.MNtype; p.mn_ak = d.ak; p.mn_psia = d.psia; p.mn_psik = d.psik; p.mn_b = d.b; p.mn_Sy = d.Sy;
  return p;
}

static void dump(const Deck &d) {
  auto vec = [](const char *name, const std::vector<double> &v) {
    std::printf("%s %zu", name, v.size());
    for (double x : v) std::printf(" %a", x);
    std::printf("\n");
  };
  auto ivec = [](const char *name, const std::vector<int32_t> &v) {
    std::printf("%s %zu", name, v.size());
    for (int x : v) std::printf(" %d", x);
    std::printf("\n");
  };
  std::printf("model %d\ndimless %d\ntimeseries %d\npiezometer %d\n", d.model, d.dimless, d.timeseries, d.piezometer);
  std::printf("M %d\nts_k %d\nts_R %d\nj0s %d %d\nnacc %d\nord %d\ntimeType %d\nMNtype %d\nzOrd %d\n", d.M, d.ts_k,
              d.ts_R, d.j0s[0], d.j0s[1], d.nacc, d.ord, d.timeType, d.MNtype, d.zOrd);
  std::printf("scalars %a %a %a %a %a %a %a %a %a %a %a %a %a %a %a\n", d.alpha, d.tol, d.kappa, d.alphaD, d.beta, d.lD,
              d.dD, d.bD, d.rDw, d.rDwobs, d.Lc, d.Tc, d.Hc, d.l, d.d);
  vec("timePar", d.timePar); vec("MoenchGamma", d.MoenchGamma);
  vec("t", d.t); vec("r", d.r); vec("z", d.z); vec("tD", d.tD); vec("rD", d.rD); vec("zD", d.zD);
  vec("j0z", d.j0z); ivec("sv", d.sv); ivec("zLay", d.zLay);
  std::printf("outfile %s\n", d.outFileName.c_str());
}

int main(int argc, char **argv) {
  std::string input = "input.dat";   // driver_io.f90:66-70
  bool fresh = false, only_dump = false, to_stdout = false, header_only = false;
  int ngpu = 0;
  for (int i = 1; i < argc; ++i) {
    std::string a = argv[i];
    if (a == "--fresh-abscissae") fresh = true;
    else if (a == "--dump") only_dump = true;
    else if (a == "--stdout") to_stdout = true;
    else if (a == "--header-only") header_only = true;   // header text to stdout, no evaluation (no device needed)
    else if (a == "--format") {                          // RFMT / HFMT of the remaining arguments (tests)
      for (int k = i + 1; k < argc; ++k) {
        const double x = std::strtod(argv[k], nullptr);
        std::printf("[%s][%s]\n", uncli::RFMT(x).c_str(), uncli::HFMT(x).c_str());
      }
      return 0;
    }
    else if (a == "--ngpu" && i + 1 < argc) ngpu = std::atoi(argv[++i]);
    else if (a == "--help" || a == "-h") {
      std::puts("usage: unconfined_cli [input.dat] [--fresh-abscissae] [--ngpu N] [--dump] [--stdout]");
      return 0;
    } else input = a;
  }
  try {
    Deck d = uncli::read_deck(input);
    fill_tables(d);
    if (d.quiet > 0)
      for (const auto &w : d.warnings) std::cout << w << "\n";
    if (only_dump) { dump(d); return 0; }
    if (header_only) {
      std::string side;
      std::cout << (d.timeseries ? uncli::timeseries_header(d, &side) : uncli::contour_header(d, &side)) << side;
      return 0;
    }

    const int32_t nt = (int32_t)d.tD.size(), nr = (int32_t)d.rD.size(), nz = (int32_t)d.zD.size();
    // The reference builds the tanh-sinh abscissae only for the first (t,r) of a run
    // (driver.f90:121-126): reproduce that by default, --fresh-abscissae uses each (t,r)'s own
    std::vector<double> scale;
    if (!fresh) scale.assign((size_t)nt * nr, d.j0z[d.sv[0] - 1] / d.rD[0]);
    std::vector<double> s((size_t)nt * nr * nz), ds(s.size());
    unc_params p = make_params(d);
    int rc = unc_eval_grid(&p, nt, d.tD.data(), d.sv.data(), nr, d.rD.data(), nz, d.zD.data(), d.zLay.data(),
                           fresh ? nullptr : scale.data(), ngpu, s.data(), ds.data());
    if (rc) throw std::runtime_error(std::string("unc_eval_grid: ") + unc_last_error());

    std::string side;
    std::string text = d.timeseries ? uncli::timeseries_header(d, &side) : uncli::contour_header(d, &side);
    text += uncli::output_rows(d, s, ds);
    std::cout << side;
    if (to_stdout) {
      std::cout << text;
    } else {
      std::ofstream f(d.outFileName, std::ios::trunc);
      if (!f) throw std::runtime_error("cannot open output file " + d.outFileName + " for writing");
      f << text;
    }
    unc_shutdown();
  } catch (const std::exception &e) {
    std::cerr << e.what() << "\n";
    return 1;
  }
  return 0;
}
