// oracle/refcheck/stubs (see Event/Event.h): host-memory stand-ins for the evk handles ShadowVoxSystem uses
// (Vendor/evk/evk.h: Buffer, Image, ImageRegion, CmdBuffer; Sources/Graphics/Graphics.h: Graphics::Transfer).
#pragma once
#include <cstdint>
#include <memory>
#include <vector>
#include <glm/glm.hpp>
#include <glm/gtc/matrix_transform.hpp>

using uint8 = uint8_t; using uint16 = uint16_t; using uint32 = uint32_t; using uint64 = uint64_t;
using int8 = int8_t; using int16 = int16_t; using int32 = int32_t; using int64 = int64_t;

struct Extent {
    uint32_t width{0}, height{0}, depth{0};
    Extent() {}
    Extent(uint32_t w, uint32_t h) : width(w), height(h), depth(1) {}
    Extent(uint32_t w, uint32_t h, uint32_t d) : width(w), height(h), depth(d) {}
};
enum class Format { R8Uint };
enum class ImageLayout { Undefined, TransferDst, ShaderReadOptimal };
struct ImageRegion { int x, y, z; uint32_t width, height, depth, mip{0}, layer{0}; };

class Image {
    Extent _e;
public:
    struct Info { Format format; Extent extent; Info(Format f, Extent e) : format(f), extent(e) {} };
    static Image Create(const Info& i) { Image im; im._e = i.extent; return im; }
    Extent getExtent() const { return _e; }
};
class Buffer {
    std::shared_ptr<std::vector<uint8_t>> _d;
public:
    static Buffer Create(uint64_t size);
    void* getData() { return _d ? _d->data() : nullptr; }
    uint64_t size() const { return _d ? _d->size() : 0; }
};
class CmdBuffer {
public:
    void barrier(Image&, ImageLayout, ImageLayout) {}
    void copy(Buffer& src, Image& dst, const std::vector<ImageRegion>& regions);
};
class Graphics {
public:
    template <typename T> static void Transfer(T callback) { CmdBuffer cmd; callback(cmd); }
};
