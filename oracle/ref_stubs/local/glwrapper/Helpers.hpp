// TEST INFRASTRUCTURE (oracle/_ref): drawing is out of scope; Trail::draw() only has to compile.
#pragma once
#include <cstddef>
namespace GL {
template<class V>
struct Span { V const* p; size_t n; };
template<class V, class R, class S, class U, class P>
void draw_with_temporary_vao(R&, S&, U const&, P, Span<V>) { }
}
