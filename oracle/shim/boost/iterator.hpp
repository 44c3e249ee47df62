#ifndef ORACLE_SHIM_BOOST_ITERATOR_HPP
#define ORACLE_SHIM_BOOST_ITERATOR_HPP
#include <iterator>
#include <boost/iterator/iterator_facade.hpp>
#endif
