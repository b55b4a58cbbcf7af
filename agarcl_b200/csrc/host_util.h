// host_util.h — error plumbing shared by the host translation units of libagarcl_b200.so.
// The reference raises C++ exceptions (EnvironmentException, BaseEnvironment.hpp:19); nothing may
// unwind across the C ABI, so every entry point records a message and returns a status instead.
#pragma once
#include "../../include/agarcl_b200.h"

int agarcl_set_error(int status, const char* fmt, ...);

// snapshot.cpp: the reference's JSON environment snapshot to / from one instance blob (host memory)
extern "C" int agarcl_snapshot_write(const agarcl_cfg* c, const agarcl_layout* L, const void* blob, const char* path);
extern "C" int agarcl_snapshot_read(const agarcl_cfg* c, const agarcl_layout* L, void* blob, const char* path, int lossless);
