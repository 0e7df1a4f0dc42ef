#ifndef ORACLE_STUB_BOOST_PRED_H_
#define ORACLE_STUB_BOOST_PRED_H_
#include <cctype>
#include <string>
namespace boost {
inline bool iequals(const std::string& a, const std::string& b) {
  if (a.size() != b.size()) return false;
  for (size_t i = 0; i < a.size(); ++i)
    if (std::tolower((unsigned char)a[i]) != std::tolower((unsigned char)b[i])) return false;
  return true;
}
}
#endif
