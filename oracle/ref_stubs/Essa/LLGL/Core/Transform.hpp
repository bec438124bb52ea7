// TEST INFRASTRUCTURE (oracle/_ref): only Trail::draw() touches this, and nothing calls Trail::draw() here.
#pragma once
#include <EssaUtil/Vector.hpp>
namespace llgl {
struct Matrix { };
struct Transform {
    Transform translate(Util::Vector3f const&) const { return *this; }
    Matrix matrix() const { return {}; }
};
}
