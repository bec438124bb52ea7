// TEST INFRASTRUCTURE (oracle/_ref): the vertex type src/Trail.hpp derives from, reduced to its position slot.
#pragma once
#include <EssaUtil/Vector.hpp>

namespace Essa::Shaders {
struct Basic {
    struct Vertex {
        Util::Point3f m_pos;
        Util::Colorf m_color;
        struct Nothing { };
        Vertex(Util::Point3f pos, Util::Colorf color, Nothing) : m_pos(pos), m_color(color) { }
        template<int I>
        Util::Point3f& value() { static_assert(I == 0); return m_pos; }
        template<int I>
        Util::Point3f const& value() const { static_assert(I == 0); return m_pos; }
    };
    struct Uniforms {
        template<class A, class B, class C>
        void set_transform(A const&, B const&, C const&) { }
    };
};
}
