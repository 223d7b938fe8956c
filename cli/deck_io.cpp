// See deck_io.hpp.  Every block cites the reference lines it restates.
#include "deck_io.hpp"

#include <cctype>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <fstream>
#include <limits>
#include <sstream>
#include <stdexcept>

namespace uncli {

const char *const kModelDescrip[7] = {"Theis", "Hantush", "Hantush w/ stor", "Moench",
                                      "Malama full pen", "Malama part pen", "Mishra/Neuman"};
const char *const kTimeDescrip[9] = {
    "step on; tpar(1) = on time; tpar(2) not used",
    "finite pulse; tpar(1:2) = on/off time",
    "infinitessimal pulse; tpar(1) = pulse location; tpar(2) not used",
    "stairs; tpar(1) = time step (Q increase by integer multiples); tpar(2) = off time",
    "rectified square wave; tpar(1) = 1/2 period of wave; tpar(2) = start time",
    "cos(omega*t); tpar(1) = omega; tpar(2) = start time",
    "rectified triangular wave; tpar(1) = 1/4 period of wave; tpar(2) = start time",
    "rectified square wave; tpar(1) = 1/2 period of wave; tpar(2) = start time",
    "piecewise constant rate (n steps); tpar(1:n)=ti; tpar(n+1)=tfinal; tpar(n+2:)=Q"};

// ---------------------------------------------------------------------------
// list-directed input
std::vector<std::string> ld_tokens(const std::string &line) {
  std::vector<std::string> out;
  std::string cur;
  auto flush = [&]() {
    if (cur.empty()) return;
    // repeat count r*c (r a positive integer)
    size_t star = cur.find('*');
    bool rep = star != std::string::npos && star > 0;
    if (rep)
      for (size_t i = 0; i < star; ++i)
        if (!std::isdigit((unsigned char)cur[i])) rep = false;
    if (rep) {
      long n = std::strtol(cur.substr(0, star).c_str(), nullptr, 10);
      for (long i = 0; i < n; ++i) out.push_back(cur.substr(star + 1));
    } else {
      out.push_back(cur);
    }
    cur.clear();
  };
  for (char c : line) {
    if (c == '/') break;  // a slash terminates a list-directed record
    if (c == ' ' || c == '\t' || c == ',' || c == '\r') flush();
    else cur.push_back(c);
  }
  flush();
  return out;
}

double ld_real(const std::string &tok) {
  std::string s = tok;
  for (char &c : s)
    if (c == 'D' || c == 'd') c = 'E';
  // Fortran also accepts an exponent without a letter (1.0-3); not used by any deck
  char *end = nullptr;
  double v = std::strtod(s.c_str(), &end);
  if (end == s.c_str() || *end != '\0') throw std::runtime_error("bad real '" + tok + "' in list-directed input");
  return v;
}

long ld_int(const std::string &tok) {
  char *end = nullptr;
  long v = std::strtol(tok.c_str(), &end, 10);
  if (end == tok.c_str() || *end != '\0') throw std::runtime_error("bad integer '" + tok + "' in list-directed input");
  return v;
}

bool ld_logical(const std::string &tok) {
  size_t i = 0;
  if (i < tok.size() && tok[i] == '.') ++i;
  if (i < tok.size()) {
    char c = (char)std::tolower((unsigned char)tok[i]);
    if (c == 't') return true;
    if (c == 'f') return false;
  }
  throw std::runtime_error("bad logical '" + tok + "' in list-directed input");
}

namespace {

// A sequential formatted file read with list-directed records: a read that needs more
// items than the current line holds continues on the next line (as Fortran does).
struct Reader {
  std::vector<std::string> lines;
  size_t pos = 0;
  std::string name;
  explicit Reader(const std::string &path) : name(path) {
    std::ifstream f(path);
    if (!f) throw std::runtime_error("ERROR opening input file " + path + " for reading");
    std::string l;
    while (std::getline(f, l)) lines.push_back(l);
  }
  std::vector<std::string> record(size_t need) {
    std::vector<std::string> t;
    while (t.size() < need) {
      if (pos >= lines.size()) throw std::runtime_error("end of file in " + name);
      auto more = ld_tokens(lines[pos++]);
      t.insert(t.end(), more.begin(), more.end());
    }
    return t;
  }
  void backspace() { if (pos > 0) --pos; }
};

std::string trim_dir(const std::string &path) {
  size_t k = path.find_last_of('/');
  return k == std::string::npos ? std::string() : path.substr(0, k + 1);
}

[[noreturn]] void stop(const std::string &msg) { throw std::runtime_error(msg); }

}  // namespace

std::vector<double> linspace(double lo, double hi, int num) {
  std::vector<double> v(num > 0 ? num : 0);
  if (num == 1) {
    v[0] = (lo + hi) / 2.0;
  } else {
    const double dx = (hi - lo) / (num - 1);
    for (int i = 1; i <= num; ++i) v[i - 1] = lo + (i - 1) * dx;
  }
  return v;
}

std::vector<double> logspace(int lo, int hi, int num) {
  std::vector<double> v = linspace((double)lo, (double)hi, num);
  for (double &x : v) x = std::pow(10.0, x);
  return v;
}

Deck read_deck(const std::string &path) {
  Deck d;
  Reader in(path);
  const std::string dir = trim_dir(path);   // data files are looked up next to the deck
  const double eps32 = (double)std::numeric_limits<float>::epsilon();   // epsilon(1.0)

  // driver_io.f90:88  quiet, model, dimless, timeseries, piezometer
  auto t = in.record(5);
  d.quiet = (int)ld_int(t[0]); d.model = (int)ld_int(t[1]);
  d.dimless = ld_logical(t[2]); d.timeseries = ld_logical(t[3]); d.piezometer = ld_logical(t[4]);
  if (d.model < 0 || d.model > 6) stop("ERROR invalid model choice " + std::to_string(d.model));
  t = in.record(1); d.Q = ld_real(t[0]);                                  // :103
  t = in.record(2); d.l = ld_real(t[0]); d.d = ld_real(t[1]);             // :107
  t = in.record(2); d.rw = ld_real(t[0]); d.rc = ld_real(t[1]);           // :110
  t = in.record(1); d.gammaSkin = ld_real(t[0]);                          // :113
  // :115-128  time behaviour, the record is read twice (backspace)
  t = in.record(1);
  d.timeType = (int)ld_int(t[0]);
  in.backspace();
  {
    int npar = 2;
    if (d.timeType <= -1) npar = -2 * (d.timeType % 100) + 1;   // mod() keeps the sign of timeType
    t = in.record(1 + (size_t)npar);
    d.timePar.resize(npar);
    for (int i = 0; i < npar; ++i) d.timePar[i] = ld_real(t[1 + i]);
  }
  t = in.record(1); d.b = ld_real(t[0]);                                  // :133
  t = in.record(2); d.Kr = ld_real(t[0]); d.kappa = ld_real(t[1]);        // :136
  t = in.record(2); d.Ss = ld_real(t[0]); d.Sy = ld_real(t[1]);           // :139
  // :143-153  beta, MoenchM [, alphas]
  t = in.record(2);
  d.beta = ld_real(t[0]); d.MoenchM = (int)ld_int(t[1]);
  in.backspace();
  if (d.model == 3 && d.MoenchM < 1) stop("ERROR: number of Moench alphas must be >= 1 for model==3");
  if (d.MoenchM < 0) d.MoenchM = 0;
  t = in.record(2 + (size_t)d.MoenchM);
  d.MoenchAlpha.resize(d.MoenchM);
  for (int i = 0; i < d.MoenchM; ++i) d.MoenchAlpha[i] = ld_real(t[2 + i]);
  // :159  Mishra/Neuman parameters
  t = in.record(7);
  d.ac = ld_real(t[0]); d.ak = ld_real(t[1]); d.psia = ld_real(t[2]); d.psik = ld_real(t[3]);
  d.usL = ld_real(t[4]); d.MNtype = (int)ld_int(t[5]); d.order = (int)ld_int(t[6]);
  if (d.MNtype == 1) {                                                    // :166-191
    if (std::fabs(d.ac - d.ak) > eps32) {
      d.warnings.push_back("WARNING1: Malama's Mishra-Neuman implementation assumes ac=ak: using ak (ignoring ac)");
      d.ac = d.ak;
    }
    if (std::fabs(d.l - d.b) > eps32) {
      d.warnings.push_back("WARNING2: Malama's Mishra-Neuman implementation assumes pumping well screen goes to bottom of formation (l=b)");
      d.l = d.b;
    }
    if (d.d > eps32) {
      d.warnings.push_back("WARNING3: Malma's Mishra-Neuman implementation assumes pumping well screen goes to top of formation (d=0)");
      d.d = 0.0;
    }
  }
  // :236-290  parameter checks
  if (d.model > 0 && (d.gammaSkin < 0.0 || d.d < 0.0 || d.l < 0.0)) stop("ERROR: negative geometry parameters (gamma, d, l)");
  if (d.b <= 0.0 || d.Kr <= 0.0 || d.Ss <= 0.0) stop("ERROR: zero or negative aquifer parametrs (b, Kr, Ss)");
  if (d.model > 2 && (d.kappa <= 0.0 || d.Sy <= 0.0)) stop("ERROR: zero or negative unconfined aquifer parameters (kappa, Sy)");
  if (d.model > 0 && d.d >= d.l) stop("ERROR: screen top/bottom (l must be > d)");
  if (d.model == 6) {
    if (d.ac < 0.0 || d.ak < 0.0 || d.usL < 0.0 || d.psia < 0.0 || d.psik < 0.0) stop("ERROR: ivalid Mishra/Neuman parameters (a_c, a_k, L, psi_a, psi_k)");
    if (d.MNtype == 2 && d.order < 3) stop("ERROR: order of Mishra/Neuman finite difference must be >=3");
  }
  if ((d.model == 4 || d.model == 5) && d.beta < 0.0) stop("ERROR: Malama linearization beta cannot be negative");
  if (d.model == 3)
    for (double a : d.MoenchAlpha)
      if (a < 0.0) stop("ERROR: Moench alphas cannot be negative");
  // :296-304  numerics
  t = in.record(3); d.M = (int)ld_int(t[0]); d.alpha = ld_real(t[1]); d.tol = ld_real(t[2]);
  t = in.record(2); d.ts_k = (int)ld_int(t[0]); d.ts_R = (int)ld_int(t[1]);
  t = in.record(4);
  d.j0s[0] = (int)ld_int(t[0]); d.j0s[1] = (int)ld_int(t[1]); d.nacc = (int)ld_int(t[2]); d.ord = (int)ld_int(t[3]);
  if (d.M < 2) stop("ERROR: deHoog number of Fourier Series terms must be >= 1");       // :308
  if (d.tol < std::numeric_limits<double>::epsilon()) {                               // :313
    d.tol = std::numeric_limits<double>::epsilon();
    d.warnings.push_back("WARNING: increased INVLAP tolerance to " + RFMT(d.tol));
  }
  if (d.ts_k - d.ts_R < 2) stop("ERROR: Tanh-Sinh k (2**k abcissa) is too low for given level of Richardson extrapolation");
  if (d.ts_R < 1) stop("ERROR: Richardson extrapolation level must be >= 1");
  if (d.j0s[0] < 1 || d.j0s[1] < 1 || d.nacc < 1 || d.ts_k < 1) stop("ERROR max/min split, # accelerated terms, and tanh-sinh k must be >= 1");
  // :341-352  locations and times
  t = in.record(2); const std::string timeFile = t[0]; const double tval = ld_real(t[1]);
  t = in.record(2); const std::string spaceFile = t[0]; const double rval = ld_real(t[1]);
  t = in.record(5);
  d.zTop = ld_real(t[0]); d.zBot = ld_real(t[1]); d.zOrd = (int)ld_int(t[2]);
  d.rwobs = ld_real(t[3]); d.sF = ld_real(t[4]);

  if (d.timeseries) {
    if (d.zTop < d.zBot) stop("ERROR: for screened observation wells top of monitoring well screen must be at or above bottom");
    if (d.zTop > d.b || d.zBot < 0.0) stop("ERROR: top of monitoring well screen must be above bottom and both between 0 and b");
    if (!d.piezometer && d.zOrd < 1) stop("ERROR: # of quadrature points at monitoring location must be > 0");
    if (d.rwobs <= 0.0) stop("ERROR: monitoring well radius must be >0");
    if (d.sF <= 0.0) stop("ERROR: monitoring well shape factor must be >0");
    if (!(rval > d.rw)) stop("ERROR: r must be > rw");
    d.r.assign(1, rval);
    if (d.piezometer) d.zOrd = 1;
    d.z = linspace(d.zBot, d.zTop, d.zOrd);                              // :404
    Reader tf(dir + timeFile);
    auto a = tf.record(2); const bool computeTimes = ld_logical(a[0]); const int numTFile = (int)ld_int(a[1]);
    a = tf.record(3);
    const int minLogT = (int)ld_int(a[0]), maxLogT = (int)ld_int(a[1]), numTComp = (int)ld_int(a[2]);
    if (computeTimes) {
      d.t = logspace(minLogT, maxLogT, numTComp);                        // :431
    } else {
      d.t.resize(numTFile);
      for (int i = 0; i < numTFile; ++i) d.t[i] = ld_real(tf.record(1)[0]);
      for (double x : d.t)
        if (x < 0.0) stop("ERROR all times must be > 0");
    }
  } else {
    if (!(tval >= 0.0)) stop("ERROR: t must be >0");
    d.t.assign(1, tval);
    Reader sf(dir + spaceFile);
    auto a = sf.record(3);
    const bool computeSpace = ld_logical(a[0]); const int numRFile = (int)ld_int(a[1]), numZFile = (int)ld_int(a[2]);
    a = sf.record(3); const double minR = ld_real(a[0]), maxR = ld_real(a[1]); const int numRComp = (int)ld_int(a[2]);
    a = sf.record(3); const double minZ = ld_real(a[0]), maxZ = ld_real(a[1]); const int numZComp = (int)ld_int(a[2]);
    if (computeSpace) {
      d.r = linspace(minR, maxR, numRComp);
      if (minZ < 0.0 || maxZ > d.b) stop("ERROR z must be in range 0<=>b");
      d.z = linspace(minZ, maxZ, numZComp);
    } else {
      a = sf.record((size_t)numRFile);
      d.r.resize(numRFile);
      for (int i = 0; i < numRFile; ++i) d.r[i] = ld_real(a[i]);
      for (double x : d.r)
        if (x < d.rw) stop("ERROR r must be >= rw");
      a = sf.record((size_t)numZFile);
      d.z.resize(numZFile);
      for (int i = 0; i < numZFile; ++i) d.z[i] = ld_real(a[i]);
      for (double x : d.z)
        if (x < 0.0 || x > d.b) stop("ERROR z must be in range 0<=>b");
    }
  }
  t = in.record(1); d.outFileName = t[0];                                 // :528

  // :534-567  characteristic and dimensionless quantities (same operation order)
  const double PI = 4.0 * std::atan(1.0);
  d.Lc = d.b;
  d.Tc = d.Lc * d.Lc / (d.Kr / d.Ss);
  d.Hc = d.Q / (4 * PI * d.Kr * d.b);
  d.malamaSigma = d.Sy / (d.Ss * d.b);
  d.alphaD = d.kappa / d.malamaSigma;
  d.lD = d.l / d.Lc; d.dD = d.d / d.Lc; d.bD = d.lD - d.dD;
  d.rDw = d.rw / d.Lc; d.rDwobs = d.rwobs / d.Lc;
  d.MoenchGamma.resize(d.MoenchM);
  for (int i = 0; i < d.MoenchM; ++i) d.MoenchGamma[i] = d.MoenchAlpha[i] * d.Lc * d.Sy / (d.kappa * d.Kr);
  d.b1 = d.psia - d.psik;
  d.zD = d.z; for (double &x : d.zD) x /= d.Lc;
  d.rD = d.r; for (double &x : d.rD) x /= d.Lc;
  d.tD = d.t; for (double &x : d.tD) x /= d.Tc;
  return d;
}

// ---------------------------------------------------------------------------
// ESw.dEe: one digit before the point, d after, exponent sign and e digits; right-justified
// in w; a field that does not fit is filled with asterisks; NaN / Infinity as gfortran.
std::string fmt_es(double x, int w, int d, int e) {
  std::string body;
  if (std::isnan(x)) {
    body = "NaN";
  } else if (std::isinf(x)) {
    body = x < 0 ? "-Infinity" : "Infinity";
    if ((int)body.size() > w) body = x < 0 ? "-Inf" : "Inf";
  } else {
    char buf[64];
    std::snprintf(buf, sizeof buf, "%.*E", d, x);   // correctly rounded decimal, like libgfortran
    std::string s(buf);
    size_t k = s.find('E');
    std::string mant = s.substr(0, k);
    int ex = std::atoi(s.c_str() + k + 1);
    char eb[16];
    std::snprintf(eb, sizeof eb, "%c%0*d", ex < 0 ? '-' : '+', e, std::abs(ex));
    if ((int)std::string(eb).size() > e + 1) return std::string((size_t)w, '*');
    body = mant + "E" + eb;
  }
  if ((int)body.size() > w) return std::string((size_t)w, '*');
  return std::string((size_t)(w - (int)body.size()), ' ') + body;
}

namespace {

std::string common_head(const Deck &d, bool ts) {
  std::ostringstream o;
  auto L = [](bool b) { return b ? "T" : "F"; };
  o << "# -*-auto-revert-*-\n";
  // EP = DP = kind 8 in the double-precision build (constants.f90:35)
  o << (ts ? "# model, EP precision :: " : "# model, EP :: ") << d.model << " " << kModelDescrip[d.model] << ", " << 8 << "\n";
  if (ts) o << "# dimensionless?, timeseries?, piezometer? :: " << L(d.dimless) << " " << L(d.timeseries) << " " << L(d.piezometer) << " \n";
  else o << "# dimensionless?, timeseries? :: " << L(d.dimless) << " " << L(d.timeseries) << " \n";
  o << "# Q (volumetric pumping rate) :: " << RFMT(d.Q) << "\n";
  o << "# b (initial sat thickness) :: " << RFMT(d.b) << "\n";
  o << "# l,d (screen bot & top) :: " << RFMT(d.l) << " " << RFMT(d.d) << " \n";
  o << "# rw,rc (well/casing radii) :: " << RFMT(d.rw) << " " << RFMT(d.rc) << " \n";
  o << (ts ? "# Kr,kappa (kappa=Kz/Kr) :: " : "# Kr,kappa (Kz/Kr) :: ") << RFMT(d.Kr) << " " << RFMT(d.kappa) << " \n";
  o << "# Ss,Sy :: " << RFMT(d.Ss) << " " << RFMT(d.Sy) << " \n";
  o << (ts ? "# gamma (dimensionless skin) :: " : "# gamma (dimless skin) :: ") << RFMT(d.gammaSkin) << "\n";
  // timeDescrip(timeType): the reference indexes the 9-entry table with timeType itself, which
  // is out of bounds for the piecewise types (< 0); entry 9 is printed for those here
  const int ti = d.timeType >= 1 && d.timeType <= 9 ? d.timeType : 9;
  o << "# pumping well time behavior :: " << d.timeType << kTimeDescrip[ti - 1];
  for (double p : d.timePar) o << RFMT(p) << " ";
  o << "\n";
  o << "# deHoog M, alpha, tol :: " << d.M << RFMT(d.alpha) << " " << RFMT(d.tol) << " \n";
  o << "# tanh-sinh: k, n extrapolation steps :: " << d.ts_k << " " << d.ts_R << " \n";
  o << "# GLquad: J0 split, n 0-accel, GL-order :: " << d.j0s[0] << " " << d.j0s[1] << " " << d.nacc << " " << d.ord << " \n";
  return o.str();
}

std::string model_lines(const Deck &d, bool ts, std::string *to_stdout) {
  std::ostringstream o;
  if (d.model == 4 || d.model == 5) {
    o << "# Malama beta linearization parameter :: " << RFMT(d.beta) << "\n";
  } else if (d.model == 3) {
    std::ostringstream s;   // written to stdout by the reference (write(*,fmt), :727-730)
    s << "# Moench Delayed Yield decay coefficients (alpha):: " << d.MoenchM;
    for (double a : d.MoenchAlpha) s << " " << RFMT(a);
    s << "\n";
    if (to_stdout) *to_stdout += s.str();
  } else if (d.model == 6) {
    o << "# Mishra/Neuman ac,ak,psia,psik,b1 ::" << RFMT(d.ac) << " " << RFMT(d.ak) << " " << RFMT(d.psia) << " "
      << RFMT(d.psik) << " " << RFMT(d.b1) << " \n";
    if (d.MNtype == 2) {
      o << (ts ? "# Mishra/Neuman vadose zone finite-difference order, finite-difference spacing ::"
               : "# Mishra/Neuman finite-difference order, finite-difference mesh spacing ::")
        << d.order << " " << RFMT(d.usL / (d.order - 1)) << "\n";
    } else if (d.MNtype == 1) {
      o << (ts ? "# NB: Malama's Mishra/Neuman implementation (1) assumes ac=ak and fully penetrating pumping well without wellbore storage\n"
               : "# NB: Malama's Mishra/Neuman implementation (1) '//&'assumes ac=ak and fully penetrating pumping well without wellbore storage\n");
    }
  }
  return o.str();
}

}  // namespace

std::string timeseries_header(const Deck &d, std::string *to_stdout) {
  std::ostringstream o;
  o << common_head(d, true);
  if (d.piezometer) {
    o << "# point obs piezometer r,rD,z,zD :: " << RFMT(d.r[0]) << " " << RFMT(d.rD[0]) << " " << RFMT(d.z[0]) << " "
      << RFMT(d.zD[0]) << " \n";
  } else {
    o << "# screened obs well r,zTop,zBot,zOrd :: " << RFMT(d.r[0]) << " " << RFMT(d.zTop) << " " << RFMT(d.zBot) << " "
      << d.zOrd << "\n";
    o << "# screened obs well rW,shape factor :: " << RFMT(d.rwobs) << " " << RFMT(d.sF) << " \n";
  }
  o << model_lines(d, true, to_stdout);
  o << "# times :: " << d.t.size() << "\n";
  o << "# characteristic length, time :: " << RFMT(d.Lc) << " " << RFMT(d.Tc) << " \n";
  const std::string title = std::string(kModelDescrip[d.model]) + "             t*dh/d(log(t))";
  if (d.dimless) {
    o << "#\n#     t_D              " << title << "\n";
  } else {
    o << "# characteristic head ::" << RFMT(d.Hc) << "\n";
    o << "#\n#     t                " << title << "\n";
  }
  o << "#---------------------------------------------------------------\n";
  return o.str();
}

std::string contour_header(const Deck &d, std::string *to_stdout) {
  std::ostringstream o;
  o << common_head(d, false);
  o << "# num r locations, rlocs :: " << d.r.size() << " ";
  for (double x : d.r) o << RFMT(x) << " ";
  o << "\n# num z locations, zlocs :: " << d.z.size() << " ";
  for (double x : d.z) o << RFMT(x) << " ";
  o << "\n# time, tD :: " << RFMT(d.t[0]) << " " << RFMT(d.tD[0]) << " \n";
  o << model_lines(d, false, to_stdout);
  const std::string title = std::string("     ") + kModelDescrip[d.model] + "          t*dh/d(log(t))";
  o << "#\n" << (d.dimless ? "#     z_D           r_D      " : "#      z            r        ") << title << "\n";
  o << "#----------------------------------------------------------------------------\n";
  return o.str();
}

std::string output_rows(const Deck &d, const std::vector<double> &totint, const std::vector<double> &totintd) {
  std::string out;
  const size_t nt = d.t.size(), nr = d.r.size(), nz = d.z.size();
  for (size_t i = 0; i < nt; ++i) {
    for (size_t k = 0; k < nr; ++k) {
      const double *f = &totint[(i * nr + k) * nz];
      const double *g = &totintd[(i * nr + k) * nz];
      if (d.timeseries) {
        double totObs, totDeriv;
        if (!d.piezometer && d.zOrd > 1) {
          // trapezoid rule across the screen exactly as written (driver.f90:236-239)
          double s1 = 0.0, s2 = 0.0;
          for (int m = 1; m < d.zOrd; ++m) { s1 += f[m]; s2 += g[m]; }   // sum(totint(2:zOrd))
          totObs = (f[0] + 2.0 * s1 + f[d.zOrd - 1]) / (2 * d.zOrd);
          totDeriv = (g[0] + 2.0 * s2 + g[d.zOrd - 1]) / (2 * d.zOrd);
        } else {
          totObs = f[0];
          totDeriv = g[0];
        }
        if (d.dimless) out += RFMT(d.tD[i]) + " " + HFMT(totObs) + " " + HFMT(totDeriv) + " \n";
        else out += RFMT(d.t[i]) + " " + HFMT(totObs * d.Hc) + " " + HFMT(totDeriv * d.Hc) + " \n";
      } else {
        for (size_t m = 0; m < nz; ++m) {
          if (d.dimless) out += RFMT(d.zD[m]) + " " + RFMT(d.rD[k]) + " " + HFMT(f[m]) + " " + HFMT(g[m]) + " \n";
          else out += RFMT(d.z[m]) + " " + RFMT(d.r[k]) + " " + HFMT(f[m] * d.Hc) + " " + HFMT(g[m] * d.Hc) + " \n";
        }
      }
    }
  }
  return out;
}

}  // namespace uncli
