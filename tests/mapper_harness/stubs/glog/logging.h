// stand-in for <glog/logging.h> (tests/mapper_harness)
#ifndef MAPPER_HARNESS_GLOG_
#define MAPPER_HARNESS_GLOG_
#include <iostream>
namespace google { inline void InitGoogleLogging(const char*) {} }
#define LOG(severity) std::cerr
#define CHECK(cond) if (!(cond)) std::cerr
#endif
