#ifndef ORACLE_SHIM_BOOST_ITERATOR_ADAPTORS_HPP
#define ORACLE_SHIM_BOOST_ITERATOR_ADAPTORS_HPP
#include <boost/iterator/iterator_facade.hpp>
#endif
