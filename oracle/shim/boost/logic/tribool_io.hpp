#ifndef ORACLE_SHIM_BOOST_TRIBOOL_IO_HPP
#define ORACLE_SHIM_BOOST_TRIBOOL_IO_HPP
#include <boost/logic/tribool.hpp>
#include <iostream>
namespace boost { namespace logic {
inline std::ostream& operator<<(std::ostream& os, tribool x) {
    if(indeterminate(x)) { return os << ((os.flags()&std::ios_base::boolalpha) ? "indeterminate" : "2"); }
    return os << static_cast<bool>(x); }
inline std::ostream& operator<<(std::ostream& os, indeterminate_keyword_t) { return os << tribool(indeterminate); }
} }
#endif
