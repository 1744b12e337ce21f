// host_c_api.h — the handle shared by the two halves of the ctypes-facing C API of the host layer.
#pragma once
#include <memory>
#include <utility>

#include "scene_api.hpp"

struct RthScene {
    rtb200::SceneSpec spec;
    std::unique_ptr<rtb200::FlatScene> flat;
    explicit RthScene(rtb200::SceneSpec s) : spec(std::move(s)) {}
};

extern "C" void rth_set_error(const char *message);  // host_c_api.cpp (thread-local, read by rth_last_error)
