// oracle/refcheck/stubs (see Event/Event.h): the members of Sources/World/Components.h:30-35,60-67 and
// Sources/Asset/VoxAsset.h:42-59 that ShadowVoxSystem touches.
#pragma once
#include "Graphics/Graphics.h"
#include <entt/entt.hpp>

class VoxAsset {
public:
    std::vector<uint8> Data;
    uint32 SizeX{0}, SizeY{0}, SizeZ{0};
    Image _Image;
    inline uint8* PixelAt(int32 X, int32 Y, int32 Z) { return Data.data() + ((size_t)X + ((size_t)Y * (size_t)SizeX) + ((size_t)Z * (size_t)SizeX * (size_t)SizeY)); }
    Image& GetImage() { return _Image; }
};
template <class T> class AssetRefT {
protected:
    T* _Asset{nullptr};
public:
    AssetRefT() {}
    AssetRefT(T* a) : _Asset(a) {}
    T* operator->() { return _Asset; }
    bool IsValid() { return _Asset != nullptr; }
};
template <class T> class AssetSlot : public AssetRefT<T> {
public:
    AssetSlot() {}
    AssetSlot(T* a) : AssetRefT<T>(a) {}
};
class PalleteAsset {};

struct VoxRenderer {
    AssetSlot<PalleteAsset> Pallete;
    AssetSlot<VoxAsset> Vox;
    int VoxSlot{-1};
    glm::vec3 Pivot{0.0f, 0.0f, 0.0f};
};
struct Transform {
    glm::mat4 PreviousWorldMatrix{1.0f};
    glm::mat4 WorldMatrix{1.0f};
    glm::mat4 Matrix{1.0f};
    glm::vec3 Position{0.0f, 0.0f, 0.0f};
    glm::vec3 Rotation{0.0f, 0.0f, 0.0f};
    glm::vec3 Scale{1.0f, 1.0f, 1.0f};
};
