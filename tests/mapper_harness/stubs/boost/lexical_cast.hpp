// stand-in for <boost/lexical_cast.hpp> (tests/mapper_harness): the stream-based conversion the reference uses for ids and paths
#ifndef MAPPER_HARNESS_BOOST_LEXICAL_CAST_
#define MAPPER_HARNESS_BOOST_LEXICAL_CAST_
#include <sstream>
#include <stdexcept>
#include <string>
namespace boost {
struct bad_lexical_cast : std::runtime_error { bad_lexical_cast() : std::runtime_error("bad lexical cast") {} };
template <typename T, typename U> T lexical_cast(const U& u) { std::stringstream ss; ss << u; T t; if (!(ss >> t)) throw bad_lexical_cast(); return t; }
}
#endif
