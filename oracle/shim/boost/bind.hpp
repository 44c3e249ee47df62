#ifndef ORACLE_SHIM_BOOST_BIND_HPP
#define ORACLE_SHIM_BOOST_BIND_HPP
#include <functional>
namespace boost { using std::bind; }
#endif
