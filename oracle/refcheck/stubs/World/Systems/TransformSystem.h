// oracle/refcheck/stubs (see Event/Event.h): Sources/World/Systems/TransformSystem.h:13,22
#pragma once
#include <entt/entt.hpp>
using entt::operator""_hs;
using Changed = entt::tag<"Changed"_hs>;
