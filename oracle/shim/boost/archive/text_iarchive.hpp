#ifndef ORACLE_SHIM_BOOST_TEXT_IARCHIVE_HPP
#define ORACLE_SHIM_BOOST_TEXT_IARCHIVE_HPP
#include <iostream>
namespace boost { namespace archive {
class text_iarchive { std::istream& _is; public: text_iarchive(std::istream& is) : _is(is) { }
    template<class T> text_iarchive& operator>>(T&) { return *this; }
    template<class T> text_iarchive& operator&(T&) { return *this; } };
} }
#endif
