// Oracle tooling only: stand-in for boost/numeric/ublas/io.hpp.
#ifndef ORACLE_SHIM_UBLAS_IO_HPP
#define ORACLE_SHIM_UBLAS_IO_HPP
#include <boost/numeric/ublas/matrix.hpp>
#include <iostream>
namespace boost { namespace numeric { namespace ublas {
template<class E> std::ostream& operator<<(std::ostream& os, const vector_expression<E>& ve) {
    const E& e=ve(); os << '[' << e.size() << "](";
    for(std::size_t i=0; i!=e.size(); ++i) { if(i) os << ','; os << e(i); } return os << ')'; }
template<class E> std::ostream& operator<<(std::ostream& os, const matrix_expression<E>& me) {
    const E& e=me(); os << '[' << e.size1() << ',' << e.size2() << "](";
    for(std::size_t i=0; i!=e.size1(); ++i) { if(i) os << ','; os << '(';
        for(std::size_t j=0; j!=e.size2(); ++j) { if(j) os << ','; os << e(i,j); } os << ')'; } return os << ')'; }
} } }
#endif
