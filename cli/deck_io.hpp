// Host side of the drop-in: the reference's deck format in, the reference's output
// layout out.  C++ restatement of driver_io.f90 (read_input :30-666, headers :668-845) and
// of the row formats of driver.f90:234-273, so that the whole product runs in an
// environment without a Fortran compiler.  The Fortran driver remains the primary caller
// (fortran/unconfined_b200_mod.f90); this is SURVEY.md section 8(f) N1.
#pragma once
#include <cstdint>
#include <string>
#include <vector>

namespace uncli {

// Everything read_input leaves behind in the reference's derived types
// (types.f90: invLaplace :31, invHankel :88, GaussLobatto :98, TanhSinh :117, well :131,
// formation :145, solution :174), flattened.
struct Deck {
  // line 1
  int quiet = 0, model = 0;
  bool dimless = false, timeseries = false, piezometer = false;
  // well
  double Q = 0, l = 0, d = 0, rw = 0, rc = 0;
  // formation
  double gammaSkin = 0, b = 0, Kr = 0, kappa = 0, Ss = 0, Sy = 0, beta = 0;
  int MoenchM = 0;
  std::vector<double> MoenchAlpha, MoenchGamma;
  double ac = 0, ak = 0, psia = 0, psik = 0, usL = 0;
  int MNtype = 0, order = 0;
  // time behaviour
  int timeType = 1;
  std::vector<double> timePar;
  // numerics
  int M = 0;
  double alpha = 0, tol = 0;
  int ts_k = 0, ts_R = 0;
  int j0s[2] = {0, 0};
  int nacc = 0, ord = 0;
  // observation geometry
  double zTop = 0, zBot = 0, rwobs = 0, sF = 0;
  int zOrd = 0;
  std::vector<double> t, r, z;
  std::string outFileName;
  // derived (driver_io.f90:531-567)
  double Lc = 0, Tc = 0, Hc = 0, malamaSigma = 0, alphaD = 0;
  double lD = 0, dD = 0, bD = 0, rDw = 0, rDwobs = 0, b1 = 0;
  std::vector<double> tD, rD, zD;
  std::vector<int32_t> zLay, sv;
  std::vector<double> j0z;
  std::vector<std::string> warnings;  // what the reference prints when quiet > 0
};

// Fortran list-directed record: tokens separated by blanks, tabs or commas; `n*v` repeats;
// reading stops once `need` items are found (the rest of the line is ignored, which is how
// the decks carry their `:: comment` tails).  Throws std::runtime_error on a bad token.
std::vector<std::string> ld_tokens(const std::string &line);
double ld_real(const std::string &tok);      // 1.0, 1.0E-3, 2.0D-2, 1d0, .5
long ld_int(const std::string &tok);
bool ld_logical(const std::string &tok);     // T F .true. .FALSE. true f...

// utility.f90:34-57
std::vector<double> linspace(double lo, double hi, int num);
std::vector<double> logspace(int lo, int hi, int num);

// read_input up to and including the non-dimensionalisation; the set-up tables that need
// the library (J0 zeros, split index, z layers: driver_io.f90:572-586,628-664) are filled
// by the caller through the C ABI (unc_j0_zeros, unc_split_index, unc_zlay).
// Throws std::runtime_error with the reference's ERROR text where read_input would `stop`.
Deck read_deck(const std::string &path);

// Fortran edit descriptors used by the reference (constants.f90:72-74)
std::string fmt_es(double x, int w, int d, int e);   // ESw.dEe
inline std::string RFMT(double x) { return fmt_es(x, 14, 7, 2); }   // 'ES14.07E2'
inline std::string HFMT(double x) { return fmt_es(x, 24, 15, 4); }  // 'ES24.15E4'

extern const char *const kModelDescrip[7];   // types.f90:194-197
extern const char *const kTimeDescrip[9];    // types.f90:68-77

// driver_io.f90:668-766 and :768-845; model 3 writes its alpha line to stdout as the
// reference does (:727-730, :814-817) -- returned in *to_stdout.
std::string timeseries_header(const Deck &d, std::string *to_stdout);
std::string contour_header(const Deck &d, std::string *to_stdout);

// driver.f90:234-272: rows for all (t, r) in the reference's loop order (t outer, r inner);
// totint/totintd are (nz,nr,nt) column-major as unc_eval_grid returns them.
std::string output_rows(const Deck &d, const std::vector<double> &totint,
                        const std::vector<double> &totintd);

}  // namespace uncli
