// TEST INFRASTRUCTURE (oracle/_ref): src/Trail.hpp includes the GUI's SimulationView.hpp for the signature of Trail::draw();
// this stand-in is linked next to the reference's Trail.* through oracle/_ref/src (see oracle/Makefile).
#pragma once
#include <Essa/LLGL/Core/Transform.hpp>
#include <cassert>
#include <utility>
#include <vector>
struct SimulationView {
    struct Camera { llgl::Matrix view_matrix() const { return {}; } };
    struct Projection { llgl::Matrix matrix() const { return {}; } };
    struct Renderer { };
    struct Shader { };
    Camera camera() const { return {}; }
    Projection projection() const { return {}; }
    Renderer& renderer() const { static Renderer r; return r; }
    Shader& basic_shader() const { static Shader s; return s; }
};
