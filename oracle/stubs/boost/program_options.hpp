#ifndef ORACLE_STUB_BOOST_PO_H_
#define ORACLE_STUB_BOOST_PO_H_
#include <stdexcept>
#include <string>
namespace boost { namespace program_options {
struct validation_error : std::runtime_error {
  enum kind_t { invalid_option_value };
  validation_error(kind_t, const std::string& m) : std::runtime_error(m) {}
};
} }
#endif
