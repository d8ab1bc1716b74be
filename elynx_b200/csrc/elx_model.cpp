// Host-side model algebra, model-string grammar, file parsers, Newick and the birth-death tree generator
// behind include/elynx_b200_model.h.  Pure C++17, no GPU.  Each block cites the reference lines it restates.
#include <algorithm>
#include <cctype>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "../../include/elynx_b200.h"
#include "../../include/elynx_b200_model.h"
#include "elx_state.hpp"

namespace {

#include "model_data.inc"

struct Err {
  std::string msg;
};
[[noreturn]] void die(const std::string &m) { throw Err{m}; }

// ------------------------------------------------------------------------------------------------
// RateMatrix.hs / SubstitutionModel.hs
// ------------------------------------------------------------------------------------------------
struct SubstModel {
  int alphabet = 0;
  int K = 0;
  std::string name;
  std::vector<double> params, pi, exch;  // exch row-major K*K
};

enum Normalize { DoNormalize, DoNotNormalize };

// totalRateWith d m = norm_1 (d <# offdiag m) (RateMatrix.hs:69-70) with Q = setDiagonal (E diag pi) (:87-103).
// The reference's SubstitutionModel.totalRate (:87-88) recovers pi from Q by `eig`; for the reversible models
// this library builds that equals the stored pi up to rounding, which is what is used here.
double total_rate(const SubstModel &sm) {
  const int K = sm.K;
  double tot = 0;
  for (int j = 0; j < K; ++j) {
    double col = 0;
    for (int i = 0; i < K; ++i)
      if (i != j) col += sm.pi[i] * (sm.exch[(size_t)i * K + j] * sm.pi[j]);
    tot += std::fabs(col);
  }
  return tot;
}

void scale_model(SubstModel &sm, double r) {  // SubstitutionModel.hs:122-125
  for (double &e : sm.exch) e *= r;
}

SubstModel substitution_model(int alphabet, const std::string &name, std::vector<double> params, Normalize nz,
                              std::vector<double> d, std::vector<double> e) {  // SubstitutionModel.hs:101-119
  double s1 = 0, s = 0;
  for (double x : d) { s1 += std::fabs(x); s += x; }
  if (!(1e-5 > std::fabs(s1 - 1.0)))  // RateMatrix.hs:53-58
    die("substitionModel: Stationary distribution does not sum to 1.0");
  SubstModel sm;
  sm.alphabet = alphabet;
  sm.K = (int)d.size();
  sm.name = name;
  sm.params = std::move(params);
  for (double &x : d) x /= s;
  sm.pi = std::move(d);
  sm.exch = std::move(e);
  if ((int)sm.exch.size() != sm.K * sm.K) die("probMatrix: Matrix is not square.");
  if (nz == DoNormalize) scale_model(sm, 1.0 / total_rate(sm));  // :129-130
  return sm;
}

int choose2(int i) { return i * (i - 1) / 2; }

std::vector<double> exch_from_list_lower(int n, const std::vector<double> &es) {  // RateMatrix.hs:150-154,219-220
  if ((int)es.size() != choose2(n)) die("exchFromListlower: the number of exchangeabilities does not match the matrix size");
  std::vector<double> m((size_t)n * n, 0.0);
  for (int i = 0; i < n; ++i)
    for (int j = 0; j < i; ++j) m[(size_t)i * n + j] = m[(size_t)j * n + i] = es[choose2(i) + j];
  return m;
}

std::vector<double> exch_from_list_upper(int n, const std::vector<double> &es) {  // RateMatrix.hs:168-172,224-225
  if ((int)es.size() != choose2(n)) die("exchFromListlower: the number of exchangeabilities does not match the matrix size");
  std::vector<double> m((size_t)n * n, 0.0);
  for (int i = 0; i < n; ++i)
    for (int j = i + 1; j < n; ++j) m[(size_t)i * n + j] = m[(size_t)j * n + i] = es[i * (n - 2) - choose2(i) + j - 1];
  return m;
}

// ------------------------------------------------------------------------------------------------
// Nucleotide.hs / AminoAcid.hs
// ------------------------------------------------------------------------------------------------
std::vector<double> ones_minus_eye(int n) {
  std::vector<double> m((size_t)n * n, 1.0);
  for (int i = 0; i < n; ++i) m[(size_t)i * n + i] = 0.0;
  return m;
}

const char kAlphaOrder[] = "ACDEFGHIKLMNPQRSTVWY";  // AminoAcid.hs:46-71
const char kPamlOrder[] = "ARNDCQEGHILKMFPSTWYV";   // AminoAcid.hs:74-99

int alpha_to_paml(int i) { return (int)(strchr(kPamlOrder, kAlphaOrder[i]) - kPamlOrder); }  // :101-107

std::vector<double> paml_to_alpha_vec(const std::vector<double> &v) {  // :118-119
  std::vector<double> o(20);
  for (int i = 0; i < 20; ++i) o[i] = v[alpha_to_paml(i)];
  return o;
}

std::vector<double> paml_to_alpha_mat(const std::vector<double> &m) {  // :126-130
  std::vector<double> o(400);
  for (int i = 0; i < 20; ++i)
    for (int j = 0; j < 20; ++j) o[(size_t)i * 20 + j] = m[(size_t)alpha_to_paml(i) * 20 + alpha_to_paml(j)];
  return o;
}

std::vector<double> norm_sum(const double *v, int n) {
  double s = 0;
  for (int i = 0; i < n; ++i) s += v[i];
  std::vector<double> o(v, v + n);
  for (double &x : o) x /= s;
  return o;
}

std::vector<double> lg_exch() { return paml_to_alpha_mat(exch_from_list_lower(20, std::vector<double>(kLgExchPamlLower, kLgExchPamlLower + 190))); }
std::vector<double> wag_exch() { return paml_to_alpha_mat(exch_from_list_lower(20, std::vector<double>(kWagExchPamlLower, kWagExchPamlLower + 190))); }
std::vector<double> lg_pi() { return paml_to_alpha_vec(norm_sum(kLgPiPaml, 20)); }    // :339-367
std::vector<double> wag_pi() { return paml_to_alpha_vec(norm_sum(kWagPiPaml, 20)); }  // :579-607

// assembleSubstitutionModel (slynx PhyloModel.hs:108-143)
SubstModel assemble(Normalize nz, const std::string &n, const std::vector<double> *ps, const std::vector<double> *sd) {
  auto need = [&](int k) {
    if ((int)sd->size() != k)
      die("Length of stationary distribution is " + std::to_string(sd->size()) + " but should be " + std::to_string(k) + ".");
  };
  const bool P = ps != nullptr, D = sd != nullptr;
  if (n == "JC" && !P && !D) return substitution_model(ELX_ALPHABET_DNA, "JC", {}, nz, std::vector<double>(4, 0.25), ones_minus_eye(4));
  if (n == "F81" && !P && D) { need(4); return substitution_model(ELX_ALPHABET_DNA, "F81", {}, nz, *sd, ones_minus_eye(4)); }
  if (n == "HKY" && P && ps->size() == 1 && D) {  // Nucleotide.hs:91-98
    need(4);
    double k = (*ps)[0];
    std::vector<double> e = {0, 1, k, 1, 1, 0, 1, k, k, 1, 0, 1, 1, k, 1, 0};
    return substitution_model(ELX_ALPHABET_DNA, "HKY", *ps, nz, *sd, e);
  }
  if (n == "GTR4" && P && D) { need(4); return substitution_model(ELX_ALPHABET_DNA, "GTR", *ps, nz, *sd, exch_from_list_upper(4, *ps)); }
  if (n == "Poisson" && !P && !D) return substitution_model(ELX_ALPHABET_PROTEIN, "Poisson", {}, nz, std::vector<double>(20, 1.0 / 20), ones_minus_eye(20));
  if (n == "Poisson-Custom" && !P && D) { need(20); return substitution_model(ELX_ALPHABET_PROTEIN, "Poisson-Custom", {}, nz, *sd, ones_minus_eye(20)); }
  if (n == "LG" && !P && !D) return substitution_model(ELX_ALPHABET_PROTEIN, "LG", {}, nz, lg_pi(), lg_exch());
  if (n == "LG-Custom" && !P && D) { need(20); return substitution_model(ELX_ALPHABET_PROTEIN, "LG-Custom", {}, nz, *sd, lg_exch()); }
  if (n == "WAG" && !P && !D) return substitution_model(ELX_ALPHABET_PROTEIN, "WAG", {}, nz, wag_pi(), wag_exch());
  if (n == "WAG-Custom" && !P && D) { need(20); return substitution_model(ELX_ALPHABET_PROTEIN, "WAG-Custom", {}, nz, *sd, wag_exch()); }
  if (n == "GTR20" && P && D) { need(20); return substitution_model(ELX_ALPHABET_PROTEIN, "GTR", *ps, nz, *sd, exch_from_list_upper(20, *ps)); }
  die("Cannot assemble substitution model. Name: \"" + n + "\"");
}

// ------------------------------------------------------------------------------------------------
// MixtureModel.hs / PhyloModel.hs
// ------------------------------------------------------------------------------------------------
struct Phylo {
  bool is_mixture = false;
  std::string name;
  std::vector<double> weights;
  std::vector<SubstModel> comps;
};

Phylo from_substitution_models(const std::string &name, Normalize nz, std::vector<double> ws, std::vector<SubstModel> sms) {
  // MixtureModel.hs:67-97
  if (ws.empty()) die("fromSubstitutionModels: No weights given.");
  if (ws.size() != sms.size()) die("fromSubstitutionModels: Number of weights and substitution models does not match.");
  for (auto &sm : sms)
    if (sm.alphabet != sms[0].alphabet) die("fromSubstitutionModels: alphabets of substitution models are not equal.");
  if (nz == DoNormalize) {
    double c = 0, sw = 0;
    for (size_t i = 0; i < sms.size(); ++i) { c += ws[i] * total_rate(sms[i]); sw += ws[i]; }
    c /= sw;
    for (auto &sm : sms) scale_model(sm, 1.0 / c);
  }
  Phylo p;
  p.is_mixture = true;
  p.name = name;
  p.weights = std::move(ws);
  p.comps = std::move(sms);
  return p;
}

// ------------------------------------------------------------------------------------------------
// model-string grammar (slynx PhyloModel.hs:73-92,145-202)
// ------------------------------------------------------------------------------------------------
struct Cursor {
  const std::string &s;
  size_t i = 0;
  bool at(char c) const { return i < s.size() && s[i] == c; }
  bool done() const { return i >= s.size(); }
};

std::string parse_name(Cursor &c) {
  size_t j = c.i;
  while (j < c.s.size() && !strchr("[]{}(),", c.s[j])) ++j;
  if (j == c.i) die("model string: expected a name at position " + std::to_string(c.i) + " in \"" + c.s + "\"");
  std::string n = c.s.substr(c.i, j - c.i);
  c.i = j;
  return n;
}

bool parse_list(Cursor &c, char open, char close, std::vector<double> &out) {
  if (!c.at(open)) return false;
  size_t j = c.s.find(close, c.i);
  if (j == std::string::npos) die(std::string("model string: missing '") + close + "'");
  std::string body = c.s.substr(c.i + 1, j - c.i - 1);
  out.clear();
  size_t p = 0;
  while (true) {
    size_t q = body.find(',', p);
    std::string tok = body.substr(p, q == std::string::npos ? std::string::npos : q - p);
    char *end = nullptr;
    double v = strtod(tok.c_str(), &end);
    if (end == tok.c_str() || *end != '\0') die("model string: cannot parse number \"" + tok + "\"");
    out.push_back(v);
    if (q == std::string::npos) break;
    p = q + 1;
  }
  c.i = j + 1;
  return true;
}

SubstModel parse_subst(Normalize nz, Cursor &c) {
  std::string n = parse_name(c);
  std::vector<double> ps, sd;
  bool hp = parse_list(c, '[', ']', ps);
  bool hd = parse_list(c, '{', '}', sd);
  if (hd) {  // PhyloModel.hs:82-92: |norm_1 f - 1| < 1e-12
    double s = 0;
    for (double x : sd) s += std::fabs(x);
    if (!(1e-12 > std::fabs(s - 1.0))) {
      char buf[96];
      snprintf(buf, sizeof buf, "Sum of stationary distribution is %.17g but should be 1.0.", s);
      die(buf);
    }
  }
  return assemble(nz, n, hp ? &ps : nullptr, hd ? &sd : nullptr);
}

Phylo cxx(int n, const std::vector<double> *mws) {  // CXXModels.hs:31-148
  const double *pi = nullptr, *w = nullptr;
  switch (n) {
    case 10: pi = kC10Pi; w = kC10W; break;
    case 20: pi = kC20Pi; w = kC20W; break;
    case 30: pi = kC30Pi; w = kC30W; break;
    case 40: pi = kC40Pi; w = kC40W; break;
    case 50: pi = kC50Pi; w = kC50W; break;
    case 60: pi = kC60Pi; w = kC60W; break;
    default: die("cxx: cannot create CXX model with " + std::to_string(n) + " components.");
  }
  std::vector<double> ws = mws ? *mws : std::vector<double>(w, w + n);
  if ((int)ws.size() != n) die("Number of weights does not match C" + std::to_string(n) + " model.");
  std::vector<SubstModel> sms;
  for (int c = 0; c < n; ++c) {
    std::vector<double> d(pi + (size_t)c * 20, pi + (size_t)(c + 1) * 20);
    double s1 = 0;
    for (double x : d) s1 += std::fabs(x);
    for (double &x : d) x /= s1;  // normalizeSD (RateMatrix.hs:61-62)
    SubstModel sm = substitution_model(ELX_ALPHABET_PROTEIN, "C" + std::to_string(n) + "; component " + std::to_string(c + 1), {},
                                       DoNotNormalize, d, ones_minus_eye(20));
    sms.push_back(std::move(sm));
  }
  return from_substitution_models("C" + std::to_string(n), DoNormalize, ws, std::move(sms));
}

Phylo get_phylo_model(const char *subst, const char *mix, bool global_norm, const std::vector<double> *mws,
                      const std::vector<std::pair<double, std::vector<double>>> *edm) {  // PhyloModel.hs:211-231
  if (!subst && !mix) die("No model was given. See help.");
  if (subst && mix) die("Both, substitution and mixture model string given; use only one.");
  if (subst) {
    if (mws) die("Weights given; but cannot be used with substitution model.");
    if (edm) die("Empirical distribution mixture model components given; but cannot be used with substitution model.");
    if (global_norm) die("Global normalization not possible for substitution models.");
    std::string s(subst);
    Cursor c{s};
    SubstModel sm = parse_subst(DoNormalize, c);
    if (!c.done()) die("model string: trailing input in \"" + s + "\"");
    Phylo p;
    p.is_mixture = false;
    p.name = sm.name;
    p.weights = {1.0};
    p.comps.push_back(std::move(sm));
    return p;
  }
  const Normalize subNz = global_norm ? DoNotNormalize : DoNormalize, mmNz = global_norm ? DoNormalize : DoNotNormalize;
  std::string m(mix);
  if (edm) {  // edmModel (:155-177)
    if (m.rfind("EDM(", 0) != 0 || m.back() != ')') die("edmModel: cannot parse \"" + m + "\"");
    std::string inner = m.substr(4, m.size() - 5);
    Cursor c{inner};
    std::string n = parse_name(c);
    std::vector<double> ps;
    bool hp = parse_list(c, '[', ']', ps);
    if (!c.done()) die("edmModel: trailing input in \"" + m + "\"");
    if (edm->empty()) die("edmModel: no EDM components given.");
    std::vector<SubstModel> sms;
    std::vector<double> ws;
    for (auto &comp : *edm) {
      sms.push_back(assemble(subNz, n, hp ? &ps : nullptr, &comp.second));
      ws.push_back(comp.first);
    }
    if (mws) ws = *mws;
    if (ws.size() != sms.size()) die("edmModel: number of substitution models and weights differs.");
    return from_substitution_models("EDM" + std::to_string(edm->size()), mmNz, ws, std::move(sms));
  }
  if (m.size() >= 2 && m[0] == 'C' && std::all_of(m.begin() + 1, m.end(), [](char ch) { return isdigit((unsigned char)ch); })) {
    if (!global_norm) die("Local normalization impossible with CXX models.");  // :180
    return cxx(atoi(m.c_str() + 1), mws);
  }
  if (!mws) die("No weights provided.");
  if (m.rfind("MIXTURE(", 0) != 0 || m.back() != ')') die("mixtureModel: cannot parse \"" + m + "\"");
  std::string inner = m.substr(8, m.size() - 9);
  Cursor c{inner};
  std::vector<SubstModel> sms;
  while (true) {
    sms.push_back(parse_subst(subNz, c));
    if (c.at(',')) { c.i++; continue; }
    break;
  }
  if (!c.done()) die("mixtureModel: trailing input in \"" + m + "\"");
  return from_substitution_models("MIXTURE", mmNz, *mws, std::move(sms));
}

// ------------------------------------------------------------------------------------------------
// GammaRateHeterogeneity.hs -- discrete gamma means, closed form (Yang 1994)
// ------------------------------------------------------------------------------------------------
// regularised lower incomplete gamma P(a, x)
double gamma_p(double a, double x) {
  if (x <= 0) return 0.0;
  if (std::isinf(x)) return 1.0;
  const double lg = std::lgamma(a);
  if (x < a + 1.0) {  // series
    double ap = a, del = 1.0 / a, sum = del;
    for (int n = 0; n < 10000; ++n) {
      ap += 1.0;
      del *= x / ap;
      sum += del;
      if (std::fabs(del) < std::fabs(sum) * 1e-17) break;
    }
    return sum * std::exp(-x + a * std::log(x) - lg);
  }
  // continued fraction (modified Lentz)
  const double tiny = 1e-300;
  double b = x + 1.0 - a, c = 1.0 / tiny, d = 1.0 / b, h = d;
  for (int i = 1; i < 10000; ++i) {
    double an = -i * (i - a);
    b += 2.0;
    d = an * d + b;
    if (std::fabs(d) < tiny) d = tiny;
    c = b + an / c;
    if (std::fabs(c) < tiny) c = tiny;
    d = 1.0 / d;
    double del = d * c;
    h *= del;
    if (std::fabs(del - 1.0) < 1e-17) break;
  }
  return 1.0 - std::exp(-x + a * std::log(x) - lg) * h;
}

// x with P(a, x) = p (bracketed Newton on the monotone CDF)
double gamma_p_inv(double a, double p) {
  if (p <= 0) return 0.0;
  if (p >= 1) return INFINITY;
  double lo = 0.0, hi = std::max(1.0, a);
  while (gamma_p(a, hi) < p) hi *= 2.0;
  double x = 0.5 * (lo + hi);
  const double lg = std::lgamma(a);
  for (int it = 0; it < 200; ++it) {
    double f = gamma_p(a, x) - p;
    if (f > 0) hi = x; else lo = x;
    double dens = std::exp(-x + (a - 1.0) * std::log(x) - lg);
    double xn = dens > 0 ? x - f / dens : 0.5 * (lo + hi);
    if (!(xn > lo && xn < hi)) xn = 0.5 * (lo + hi);
    if (std::fabs(xn - x) <= 1e-16 * std::fabs(x)) { x = xn; break; }
    x = xn;
  }
  return x;
}

std::vector<double> gamma_means(int n, double alpha) {  // getMeans (:88-107)
  if (n <= 0) die("getMeans: Number of rate categories is zero or negative.");
  if (!(alpha > 0)) die("getMeans: shape parameter must be positive.");
  std::vector<double> cdf(n + 1, 0.0), out(n);
  for (int i = 1; i < n; ++i) {
    double q = gamma_p_inv(alpha, (double)i / n) / alpha;  // quantile of Gamma(shape alpha, scale 1/alpha)
    cdf[i] = gamma_p(alpha + 1.0, alpha * q);
  }
  cdf[n] = 1.0;
  for (int i = 0; i < n; ++i) out[i] = n * (cdf[i + 1] - cdf[i]);
  return out;
}

void expand_gamma(Phylo &p, int n, double alpha) {  // expand (:47-83)
  std::vector<double> means = gamma_means(n, alpha);
  Phylo out;
  out.is_mixture = true;
  out.name = p.name + " with discrete gamma rate heterogeneity";
  for (int k = 0; k < n; ++k)  // category-major: V.concatMap components over the scaled copies
    for (size_t c = 0; c < p.comps.size(); ++c) {
      SubstModel sm = p.comps[c];
      scale_model(sm, means[k]);
      sm.name += "; gamma rate category " + std::to_string(k + 1);
      out.comps.push_back(std::move(sm));
      out.weights.push_back(p.is_mixture ? p.weights[c] : 1.0);
    }
  p = std::move(out);
}

// ------------------------------------------------------------------------------------------------
// file parsers
// ------------------------------------------------------------------------------------------------
std::vector<std::string> split_lines(const std::string &s) {
  std::vector<std::string> out;
  size_t p = 0;
  while (p <= s.size()) {
    size_t q = s.find('\n', p);
    std::string ln = s.substr(p, q == std::string::npos ? std::string::npos : q - p);
    if (!ln.empty() && ln.back() == '\r') ln.pop_back();
    out.push_back(ln);
    if (q == std::string::npos) break;
    p = q + 1;
  }
  return out;
}

std::vector<double> parse_doubles(const std::string &ln) {
  std::vector<double> xs;
  const char *p = ln.c_str();
  while (true) {
    while (*p == ' ' || *p == '\t') ++p;
    if (!*p) break;
    char *end = nullptr;
    double v = strtod(p, &end);
    if (end == p) die("cannot parse number in line \"" + ln + "\"");
    xs.push_back(v);
    p = end;
  }
  return xs;
}

bool blank(const std::string &s) { return std::all_of(s.begin(), s.end(), [](char c) { return isspace((unsigned char)c); }); }

using Edm = std::vector<std::pair<double, std::vector<double>>>;

Edm parse_edm(const std::string &text) {  // EDMModelPhylobayes.hs:33-64
  std::vector<std::string> lines = split_lines(text);
  size_t li = 0;
  while (li < lines.size() && blank(lines[li])) ++li;
  if (li == lines.size()) die("phylobayes: headerLine");
  const std::string &hdr = lines[li++];
  char *end = nullptr;
  long n = strtol(hdr.c_str(), &end, 10);
  if (end == hdr.c_str()) die("phylobayes: headerLine");
  std::string rest(end);
  size_t a = rest.find_first_not_of(" \t");
  rest = a == std::string::npos ? "" : rest.substr(a);
  while (!rest.empty() && isspace((unsigned char)rest.back())) rest.pop_back();
  if (rest != "A C D E F G H I K L M N P Q R S T V W Y" && rest != "A C G T") die("phylobayes: headerLine");
  while (li < lines.size() && blank(lines[li])) ++li;
  if (li == lines.size()) die("phylobayes: kComponentsLine");
  long k = strtol(lines[li++].c_str(), &end, 10);
  Edm out;
  for (; li < lines.size(); ++li) {
    if (blank(lines[li])) continue;
    std::vector<double> xs = parse_doubles(lines[li]);
    if ((long)xs.size() - 1 != n) die("Did not find correct number of entries.");
    out.emplace_back(xs[0], std::vector<double>(xs.begin() + 1, xs.end()));
  }
  if ((long)out.size() != k) die("phylobayes: number of components does not match the header");
  return out;
}

Edm parse_siteprofiles(const std::string &text) {  // SiteprofilesPhylobayes.hs:33-62
  std::vector<std::string> lines = split_lines(text);
  Edm out;
  for (size_t li = 1; li < lines.size(); ++li) {
    if (blank(lines[li])) continue;
    std::vector<double> xs = parse_doubles(lines[li]);
    if (xs.size() < 2) die("siteprofiles: dataLine");
    out.emplace_back(1.0, std::vector<double>(xs.begin() + 1, xs.end()));
  }
  for (auto &c : out)
    if (c.second.size() != out[0].second.size()) die("The site profiles have a different number of entries.");
  return out;
}

int edm_to_c(const Edm &e, int *n_components, int *k, double **weights, double **pis) {
  *n_components = (int)e.size();
  *k = e.empty() ? 0 : (int)e[0].second.size();
  *weights = (double *)malloc(sizeof(double) * std::max<size_t>(1, e.size()));
  *pis = (double *)malloc(sizeof(double) * std::max<size_t>(1, e.size() * (size_t)*k));
  for (size_t i = 0; i < e.size(); ++i) {
    (*weights)[i] = e[i].first;
    for (int j = 0; j < *k; ++j) (*pis)[i * (size_t)*k + j] = e[i].second[j];
  }
  return 0;
}

}  // namespace

// ------------------------------------------------------------------------------------------------
// trees
// ------------------------------------------------------------------------------------------------
struct elx_flat_tree {
  std::vector<int32_t> parent, leaf_row;
  std::vector<double> brlen;
  std::vector<std::string> names;
  int n_leaves = 0;
};

struct elx_phylo_model {
  Phylo p;
};

namespace {

// newick Standard (Newick.hs:90-201) + toLengthTree (Phylogeny.hs:374-381)
elx_flat_tree *parse_newick(const std::string &s) {
  auto t = new elx_flat_tree();
  std::vector<char> has_len;
  size_t pos = 0;
  auto skip_ws = [&]() { while (pos < s.size() && isspace((unsigned char)s[pos])) ++pos; };
  auto fail_at = [&](const std::string &what) { delete t; die("newick: " + what + " at position " + std::to_string(pos)); };
  auto new_node = [&](int par) {
    t->parent.push_back(par);
    t->brlen.push_back(0.0);
    t->leaf_row.push_back(-1);
    t->names.emplace_back();
    has_len.push_back(0);
    return (int)t->parent.size() - 1;
  };
  auto name_and_phylo = [&](int node) {
    skip_ws();
    if (pos < s.size() && s[pos] == '\'') {
      size_t j = s.find('\'', pos + 1);
      if (j == std::string::npos) fail_at("unterminated quoted name");
      t->names[node] = s.substr(pos + 1, j - pos - 1);
      pos = j + 1;
    } else {
      size_t j = pos;
      while (j < s.size() && !strchr(" :;()[],", s[j]) && !isspace((unsigned char)s[j])) ++j;
      t->names[node] = s.substr(pos, j - pos);
      pos = j;
    }
    size_t save = pos;
    skip_ws();
    if (pos < s.size() && s[pos] == ':') {
      ++pos;
      skip_ws();
      char *end = nullptr;
      double v = strtod(s.c_str() + pos, &end);
      if (end == s.c_str() + pos) fail_at("branchLengthSimple; double");
      if (v < 0) fail_at("negative branch length");
      t->brlen[node] = v;
      has_len[node] = 1;
      pos = (size_t)(end - s.c_str());
    } else {
      pos = save;
    }
    if (pos < s.size() && s[pos] == '[') {  // branch support, ignored
      size_t j = s.find(']', pos);
      if (j == std::string::npos) fail_at("branchSupportEnd");
      pos = j + 1;
    }
  };
  skip_ws();
  std::vector<int> stack;  // open internal nodes
  int cur_parent = -1;
  bool expect_tree = true;
  while (true) {
    if (expect_tree) {
      if (pos < s.size() && s[pos] == '(') {
        int me = new_node(cur_parent);
        stack.push_back(me);
        cur_parent = me;
        ++pos;
        continue;  // first child follows immediately (no space skipping: `char '(' *> tree`)
      }
      int me = new_node(cur_parent);
      t->leaf_row[me] = t->n_leaves++;
      name_and_phylo(me);
      expect_tree = false;
    } else {
      if (stack.empty()) break;
      if (pos < s.size() && s[pos] == ',') {
        ++pos;
        skip_ws();
        expect_tree = true;
      } else if (pos < s.size() && s[pos] == ')') {
        ++pos;
        int me = stack.back();
        stack.pop_back();
        cur_parent = stack.empty() ? -1 : stack.back();
        name_and_phylo(me);
      } else {
        fail_at("expected ',' or ')'");
      }
    }
  }
  if (pos >= s.size() || s[pos] != ';') fail_at("expected ';'");
  for (size_t n = 1; n < has_len.size(); ++n)
    if (!has_len[n]) { delete t; die("toLengthTree: Length unavailable for some branches."); }
  return t;
}

struct SplitMix64 {
  uint64_t s;
  uint64_t next() {
    uint64_t z = (s += 0x9E3779B97F4A7C15ull);
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
  }
  double uniform() { return ((double)(next() >> 11) + 0.5) * 0x1p-53; }  // (0,1)
};

// PointProcess.hs: simulateRandom (:119-148) + simulateOrigin (:151-177) + toReconstructedTree (:271-317)
elx_flat_tree *birth_death(int n, double l, double m, uint64_t seed) {
  if (n < 2) die("Number of samples needs to be two or larger.");
  if (!(l > 0)) die("Birth rate needs to be positive.");
  if (m > l) die("simulateRandom: Please specify height if mu > lambda.");
  SplitMix64 g{seed};
  std::vector<double> vs(n - 1);
  double origin;
  const double d = l - m;
  if (std::fabs(d) <= 1e-12) {
    // critical process with unknown origin (BirthDeathCriticalNoTime): quantile p / (l (1 - p))
    for (double &v : vs) { double p = g.uniform(); v = p / (l * (1.0 - p)); }
    origin = *std::max_element(vs.begin(), vs.end());
  } else {
    if (std::fabs(d) <= 1e-5) die("birth-death: near-critical process (|lambda - mu| <= 1e-5) is not supported");
    {  // TimeOfOrigin.quantile (TimeOfOrigin.hs:79-92)
      double p = g.uniform(), pn = std::pow(p, 1.0 / n);
      origin = -1.0 / d * std::log(l * (1.0 - pn) / (l - pn * m));
    }
    const double t2 = (l - m * std::exp(-d * origin)) / (1.0 - std::exp(-d * origin));
    for (double &v : vs) {  // BirthDeath.quantile (BirthDeath.hs:80-92)
      double p = g.uniform();
      v = (-1.0 / d) * std::log((1.0 - p * l / t2) / (1.0 - p * m / t2));
    }
  }
  // max-Cartesian tree over the n-1 merge heights: internal node k sits between leaves k and k+1
  const int M = n - 1;
  std::vector<int> left(M, -1), right(M, -1), stack;
  for (int k = 0; k < M; ++k) {
    int last = -1;
    while (!stack.empty() && vs[stack.back()] < vs[k]) { last = stack.back(); stack.pop_back(); }
    left[k] = last;
    if (!stack.empty()) right[stack.back()] = k;
    stack.push_back(k);
  }
  int root = stack.front();
  auto t = new elx_flat_tree();
  // preorder emission; children of internal node k: left subtree (or leaf k), right subtree (or leaf k+1)
  struct Item { int kind, id, parent; double parent_h; };  // kind 0 = internal, 1 = leaf
  std::vector<Item> todo;
  todo.push_back({0, root, -1, origin});
  while (!todo.empty()) {
    Item it = todo.back();
    todo.pop_back();
    int me = (int)t->parent.size();
    t->parent.push_back(it.parent);
    if (it.kind == 1) {
      t->brlen.push_back(it.parent_h);
      t->leaf_row.push_back(it.id);
      t->names.push_back(std::to_string(it.id));
      t->n_leaves++;
    } else {
      double h = vs[it.id];
      t->brlen.push_back(it.parent_h - h);
      t->leaf_row.push_back(-1);
      t->names.push_back("0");
      // push right first so that the left child is emitted next (preorder, left to right)
      if (right[it.id] >= 0) todo.push_back({0, right[it.id], me, h}); else todo.push_back({1, it.id + 1, me, h});
      if (left[it.id] >= 0) todo.push_back({0, left[it.id], me, h}); else todo.push_back({1, it.id, me, h});
    }
  }
  return t;
}

// Haskell `show` for Double / ByteString.Builder.doubleDec: shortest round-trip digits, scientific outside [0.1, 10^7)
std::string double_dec(double x) {
  if (x == 0) return "0.0";
  char buf[64];
  int prec = 1;
  for (; prec <= 17; ++prec) {
    snprintf(buf, sizeof buf, "%.*e", prec - 1, x);
    if (strtod(buf, nullptr) == x) break;
  }
  std::string s(buf);
  size_t epos = s.find('e');
  std::string mant = s.substr(0, epos);
  int e10 = atoi(s.c_str() + epos + 1);
  bool neg = mant[0] == '-';
  if (neg) mant = mant.substr(1);
  std::string digits;
  for (char c : mant) if (isdigit((unsigned char)c)) digits.push_back(c);
  while (digits.size() > 1 && digits.back() == '0') digits.pop_back();
  std::string out = neg ? "-" : "";
  double ax = std::fabs(x);
  if (ax >= 0.1 && ax < 1e7) {
    if (e10 >= 0) {
      std::string ip = digits.substr(0, std::min(digits.size(), (size_t)e10 + 1));
      ip.append((size_t)e10 + 1 - ip.size(), '0');
      std::string fp = digits.size() > (size_t)e10 + 1 ? digits.substr((size_t)e10 + 1) : "0";
      out += ip + "." + fp;
    } else {
      out += "0." + std::string((size_t)(-e10 - 1), '0') + digits;
    }
  } else {
    out += digits.substr(0, 1) + "." + (digits.size() > 1 ? digits.substr(1) : "0") + "e" + std::to_string(e10);
  }
  return out;
}

std::string to_newick(const elx_flat_tree &t) {  // Export/Newick.hs:49-65: (forest)name:length
  const int N = (int)t.parent.size();
  std::vector<std::vector<int>> kids(N);
  for (int n = 1; n < N; ++n) kids[t.parent[n]].push_back(n);
  std::string out;
  struct Fr { int node; size_t next; };
  std::vector<Fr> st;
  st.push_back({0, 0});
  if (!kids[0].empty()) out += "(";
  while (!st.empty()) {
    Fr &f = st.back();
    if (f.next < kids[f.node].size()) {
      if (f.next > 0) out += ",";
      int c = kids[f.node][f.next++];
      if (!kids[c].empty()) out += "(";
      st.push_back({c, 0});
    } else {
      if (!kids[f.node].empty()) out += ")";
      out += t.names[f.node] + ":" + double_dec(t.brlen[f.node]);
      st.pop_back();
    }
  }
  out += ";";
  return out;
}

template <typename F>
int guarded(F &&f) {
  try {
    f();
    return 0;
  } catch (const Err &e) {
    return elx::fail(nullptr, ELX_E_INVALID, e.msg);
  } catch (const std::exception &e) {
    return elx::fail(nullptr, ELX_E_INVALID, e.what());
  }
}

}  // namespace

extern "C" {

int elx_phylo_model_parse(const char *subst, const char *mix, int global_norm, const double *weights, int n_weights,
                          const double *edm_weights, const double *edm_pis, int n_edm, int edm_k, elx_phylo_model **out) {
  if (!out) return elx::fail(nullptr, ELX_E_INVALID, "elx_phylo_model_parse: out is NULL");
  *out = nullptr;
  return guarded([&]() {
    std::vector<double> ws;
    if (weights) ws.assign(weights, weights + n_weights);
    Edm edm;
    if (edm_pis)
      for (int i = 0; i < n_edm; ++i)
        edm.emplace_back(edm_weights ? edm_weights[i] : 1.0, std::vector<double>(edm_pis + (size_t)i * edm_k, edm_pis + (size_t)(i + 1) * edm_k));
    Phylo p = get_phylo_model(subst, mix, global_norm != 0, weights ? &ws : nullptr, edm_pis ? &edm : nullptr);
    *out = new elx_phylo_model{std::move(p)};
  });
}

int elx_phylo_model_expand_gamma(elx_phylo_model *m, int n_categories, double alpha) {
  if (!m) return elx::fail(nullptr, ELX_E_INVALID, "elx_phylo_model_expand_gamma: NULL model");
  return guarded([&]() { expand_gamma(m->p, n_categories, alpha); });
}

void elx_phylo_model_free(elx_phylo_model *m) { delete m; }

int elx_phylo_model_info(const elx_phylo_model *m, int *n_states, int *n_components, int *is_mixture, int *alphabet) {
  if (!m) return elx::fail(nullptr, ELX_E_INVALID, "elx_phylo_model_info: NULL model");
  if (n_states) *n_states = m->p.comps[0].K;
  if (n_components) *n_components = (int)m->p.comps.size();
  if (is_mixture) *is_mixture = m->p.is_mixture ? 1 : 0;
  if (alphabet) *alphabet = m->p.comps[0].alphabet;
  return 0;
}

int elx_phylo_model_arrays(const elx_phylo_model *m, double *weights, double *pi, double *exch) {
  if (!m) return elx::fail(nullptr, ELX_E_INVALID, "elx_phylo_model_arrays: NULL model");
  const int K = m->p.comps[0].K;
  for (size_t c = 0; c < m->p.comps.size(); ++c) {
    if (weights) weights[c] = m->p.weights[c];
    if (pi) memcpy(pi + c * K, m->p.comps[c].pi.data(), sizeof(double) * K);
    if (exch) memcpy(exch + c * K * K, m->p.comps[c].exch.data(), sizeof(double) * K * K);
  }
  return 0;
}

const char *elx_phylo_model_name(const elx_phylo_model *m, int c) {
  if (!m) return "";
  if (c < 0 || c >= (int)m->p.comps.size()) return m->p.name.c_str();
  return m->p.comps[c].name.c_str();
}

int elx_gamma_means(int n, double alpha, double *out) {
  if (!out) return elx::fail(nullptr, ELX_E_INVALID, "elx_gamma_means: out is NULL");
  return guarded([&]() {
    std::vector<double> v = gamma_means(n, alpha);
    memcpy(out, v.data(), sizeof(double) * n);
  });
}

int elx_parse_edm_phylobayes(const char *text, int64_t len, int *n_components, int *k, double **weights, double **pis) {
  if (!text || !n_components || !k || !weights || !pis) return elx::fail(nullptr, ELX_E_INVALID, "elx_parse_edm_phylobayes: NULL argument");
  return guarded([&]() { edm_to_c(parse_edm(std::string(text, (size_t)len)), n_components, k, weights, pis); });
}

int elx_parse_siteprofiles_phylobayes(const char *text, int64_t len, int *n_components, int *k, double **weights, double **pis) {
  if (!text || !n_components || !k || !weights || !pis) return elx::fail(nullptr, ELX_E_INVALID, "elx_parse_siteprofiles_phylobayes: NULL argument");
  return guarded([&]() { edm_to_c(parse_siteprofiles(std::string(text, (size_t)len)), n_components, k, weights, pis); });
}

void elx_free(void *p) { free(p); }

int elx_newick_parse(const char *text, int64_t len, elx_flat_tree **out) {
  if (!text || !out) return elx::fail(nullptr, ELX_E_INVALID, "elx_newick_parse: NULL argument");
  *out = nullptr;
  return guarded([&]() { *out = parse_newick(std::string(text, (size_t)len)); });
}

int elx_birth_death_tree(int n_leaves, double lambda, double mu, uint64_t seed, elx_flat_tree **out) {
  if (!out) return elx::fail(nullptr, ELX_E_INVALID, "elx_birth_death_tree: out is NULL");
  *out = nullptr;
  return guarded([&]() { *out = birth_death(n_leaves, lambda, mu, seed); });
}

void elx_flat_tree_free(elx_flat_tree *t) { delete t; }

int elx_flat_tree_size(const elx_flat_tree *t, int *n_nodes, int *n_leaves) {
  if (!t) return elx::fail(nullptr, ELX_E_INVALID, "elx_flat_tree_size: NULL tree");
  if (n_nodes) *n_nodes = (int)t->parent.size();
  if (n_leaves) *n_leaves = t->n_leaves;
  return 0;
}

int elx_flat_tree_arrays(const elx_flat_tree *t, int32_t *parent, double *brlen, int32_t *leaf_row) {
  if (!t) return elx::fail(nullptr, ELX_E_INVALID, "elx_flat_tree_arrays: NULL tree");
  const size_t n = t->parent.size();
  if (parent) memcpy(parent, t->parent.data(), sizeof(int32_t) * n);
  if (brlen) memcpy(brlen, t->brlen.data(), sizeof(double) * n);
  if (leaf_row) memcpy(leaf_row, t->leaf_row.data(), sizeof(int32_t) * n);
  return 0;
}

const char *elx_flat_tree_name(const elx_flat_tree *t, int node) {
  if (!t || node < 0 || node >= (int)t->names.size()) return "";
  return t->names[node].c_str();
}

int elx_flat_tree_to_newick(const elx_flat_tree *t, char **out) {
  if (!t || !out) return elx::fail(nullptr, ELX_E_INVALID, "elx_flat_tree_to_newick: NULL argument");
  std::string s = to_newick(*t);
  *out = (char *)malloc(s.size() + 1);
  memcpy(*out, s.c_str(), s.size() + 1);
  return 0;
}

}  // extern "C"
