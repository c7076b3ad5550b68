#include "ptree.hpp"
