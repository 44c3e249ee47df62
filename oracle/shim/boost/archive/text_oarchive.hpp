// Oracle tooling only: inert stand-in (serialization is not on the reachability path).
#ifndef ORACLE_SHIM_BOOST_TEXT_OARCHIVE_HPP
#define ORACLE_SHIM_BOOST_TEXT_OARCHIVE_HPP
#include <iostream>
namespace boost { namespace archive {
class text_oarchive { std::ostream& _os; public: text_oarchive(std::ostream& os) : _os(os) { }
    template<class T> text_oarchive& operator<<(const T&) { return *this; }
    template<class T> text_oarchive& operator&(const T&) { return *this; } };
} }
#endif
