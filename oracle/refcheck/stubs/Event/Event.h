// oracle/refcheck/stubs -- TEST INFRASTRUCTURE: minimal stand-ins for the engine headers that
// Sources/World/Systems/ShadowVoxSystem.{h,cpp} include, so that the reference's voxeliser compiles
// FROM WHERE IT LIES (with the reference's vendored entt and glm) without Vulkan.  Interface names only.
#pragma once
struct Event {};
