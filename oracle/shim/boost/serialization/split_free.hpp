// oracle/shim: forwards to the single Boost stand-in (TEST INFRASTRUCTURE ONLY, see crf_boost_shim.hpp)
#include <boost/crf_boost_shim.hpp>
