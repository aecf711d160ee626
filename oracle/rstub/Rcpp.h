/* oracle/rstub/Rcpp.h -- TEST INFRASTRUCTURE (see R.h).
 * Declaration-level stand-in for the absent Rcpp headers, just enough for the reference's src/vertex_tfce.cpp to
 * compile from where it lies: it uses std::vector arguments, Rcpp::stop and Rcpp::checkUserInterrupt only. */
#ifndef RSTUB_RCPP_H
#define RSTUB_RCPP_H
#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <numeric>
#include <stdexcept>
#include <string>
#include <vector>
namespace Rcpp {
inline void stop(const std::string &msg) { throw std::runtime_error(msg); }
inline void checkUserInterrupt() {}
}  // namespace Rcpp
#endif
