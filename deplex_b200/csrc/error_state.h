// error_state.h -- per-thread message for failures that have no handle (dpx_create, dpx_config_load_ini).
#pragma once
#include <string>

namespace dpx {
void set_thread_error(const std::string& msg);
const char* thread_error();
}  // namespace dpx
