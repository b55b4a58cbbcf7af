// host_util.h — error plumbing shared by the host translation units of libagarcl_b200.so.
// The reference raises C++ exceptions (EnvironmentException, BaseEnvironment.hpp:19); nothing may
// unwind across the C ABI, so every entry point records a message and returns a status instead.
#pragma once
#include "../../include/agarcl_b200.h"

int agarcl_set_error(int status, const char* fmt, ...);
