// Oracle tooling only: stand-in for boost/shared_ptr.hpp (Boost is absent from this image).
// Maps the subset of boost smart pointers the reference uses onto the C++11 standard library.
#ifndef ORACLE_SHIM_BOOST_SHARED_PTR_HPP
#define ORACLE_SHIM_BOOST_SHARED_PTR_HPP
#include <memory>
namespace boost {
using std::shared_ptr;
using std::weak_ptr;
using std::make_shared;
using std::dynamic_pointer_cast;
using std::static_pointer_cast;
using std::const_pointer_cast;
using std::enable_shared_from_this;
}
#endif
