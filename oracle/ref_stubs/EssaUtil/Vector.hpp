// TEST INFRASTRUCTURE (oracle/_ref): stand-in for the un-vendored EssaUtil/Vector.hpp (EssaGUI @ 6bc4531), just wide enough
// for the reference's own src/History.cpp and src/Trail.cpp to compile unchanged.  What those two files pin is their CONTROL
// FLOW (ring indices, append offset / length bookkeeping, the decimation rule); the vector arithmetic below encodes the
// assumptions DESIGN.md section 3 lists as unverifiable offline: (U1) vector / scalar = three IEEE divisions,
// (U2) AU = 149,597,870,700 m, (U3) Point3f / double evaluated in double, then narrowed.
#pragma once
#include <cmath>
#include <cstdint>
#include <ostream>

namespace Util {

template<class T>
struct V3 {
    T m_x {}, m_y {}, m_z {};
    V3() = default;
    V3(T x, T y, T z) : m_x(x), m_y(y), m_z(z) { }
    T& x() { return m_x; }
    T& y() { return m_y; }
    T& z() { return m_z; }
    T x() const { return m_x; }
    T y() const { return m_y; }
    T z() const { return m_z; }
    template<class U>
    V3<U> cast() const { return { static_cast<U>(m_x), static_cast<U>(m_y), static_cast<U>(m_z) }; }
    V3 operator+(V3 const& o) const { return { static_cast<T>(m_x + o.m_x), static_cast<T>(m_y + o.m_y), static_cast<T>(m_z + o.m_z) }; }
    V3 operator-(V3 const& o) const { return { static_cast<T>(m_x - o.m_x), static_cast<T>(m_y - o.m_y), static_cast<T>(m_z - o.m_z) }; }
    // (U1) / (U3): the quotient is formed in double and narrowed to T
    V3 operator/(double s) const {
        return { static_cast<T>(static_cast<double>(m_x) / s), static_cast<T>(static_cast<double>(m_y) / s), static_cast<T>(static_cast<double>(m_z) / s) };
    }
    bool operator==(V3 const& o) const { return m_x == o.m_x && m_y == o.m_y && m_z == o.m_z; }
};

using DeprecatedVector3d = V3<double>;
using Vector3d = V3<double>;
using Point3d = V3<double>;
using Vector3f = V3<float>;
using Point3f = V3<float>;

struct Color {
    uint8_t r = 255, g = 255, b = 255, a = 255;
};
struct Colorf {
    float r = 1, g = 1, b = 1, a = 1;
    Colorf() = default;
    Colorf(Color const& c) : r(c.r / 255.f), g(c.g / 255.f), b(c.b / 255.f), a(c.a / 255.f) { }
};

namespace Constants {
inline constexpr double AU = 149597870700.0; // (U2)
}

} // namespace Util
