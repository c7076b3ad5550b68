// stand-in for <boost/chrono.hpp> (tests/mapper_harness): std::chrono under the boost names util/timer.h uses
#ifndef MAPPER_HARNESS_BOOST_CHRONO_
#define MAPPER_HARNESS_BOOST_CHRONO_
#include <chrono>
namespace boost { namespace chrono { using namespace std::chrono; } }
#endif
