// primitive3d_b200/csrc/p3d_error.h -- the per-thread message behind p3d_last_error().
#pragma once
#include <string>

#include "../../include/prim3d_b200.h"

namespace p3d {
p3d_status set_error(p3d_status st, const std::string &msg);  // stores msg, returns st
}
