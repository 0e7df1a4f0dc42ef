#ifndef ORACLE_STUB_BOOST_REPL_H_
#define ORACLE_STUB_BOOST_REPL_H_
#include <string>
namespace boost {
inline std::string replace_all_copy(std::string s, const std::string& from, const std::string& to) {
  size_t pos = 0;
  while ((pos = s.find(from, pos)) != std::string::npos) { s.replace(pos, from.size(), to); pos += to.size(); }
  return s;
}
}
#endif
