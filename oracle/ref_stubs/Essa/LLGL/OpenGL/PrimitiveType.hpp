#pragma once
namespace llgl { enum class PrimitiveType { LineStrip }; }
