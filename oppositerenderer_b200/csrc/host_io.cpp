// host_io.cpp — host-only I/O either side of the hot path:
//   * display normalisation + BMP writer: Gui/gui/RenderWidget.cpp:187-199 (shader: pow(color * invIterations, 1/gamma)) and
//     :243-317 (saveImageAsBMP: 54-byte header, 24-bit BGR, rows in buffer order, no row padding — kept);
//   * texture image loading: RenderEngine/util/Image.cpp:17-140 — uncompressed TGA types 2 / 3 through libtga with TGA_RGB
//     (BGR -> RGB for 24/32-bit), no vertical flip, alpha byte = 1 (Image.cpp:100-110).  The reference's Qt branch for other
//     formats never sets width/height (Image.cpp:39-60), i.e. yields a 0x0 texture; here 24/32-bit BMP files load properly
//     instead (documented deviation).
// No CUDA here.
#include "../../include/opr.h"
#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>

namespace {

thread_local std::string g_ioError;
int io_fail(const std::string& m) { g_ioError = m; return 1; }

#pragma pack(push, 1)
struct BmpHeader {   // RenderWidget.cpp:243-261 (follows the two bytes "BM")
    uint32_t mFileSize, mReserved01, mDataOffset, mHeaderSize;
    int32_t mWidth, mHeight;
    int16_t mColorPlates, mBitsPerPixel;
    uint32_t mCompression, mImageSize, mHorizRes, mVertRes, mPaletteColors, mImportantColors;
};
#pragma pack(pop)
static_assert(sizeof(BmpHeader) == 52, "BMP header layout");

inline unsigned char to_byte(float v) { return (unsigned char)(std::min(255.f, std::max(0.f, v))); }

bool read_file(const char* path, std::vector<unsigned char>& out)
{
    FILE* f = fopen(path, "rb");
    if (!f) return false;
    fseek(f, 0, SEEK_END); long n = ftell(f); fseek(f, 0, SEEK_SET);
    out.resize(n > 0 ? (size_t)n : 0);
    size_t got = out.empty() ? 0 : fread(out.data(), 1, out.size(), f);
    fclose(f);
    return got == out.size();
}

bool ends_with(const std::string& s, const char* suf)
{
    std::string l = s; std::transform(l.begin(), l.end(), l.begin(), ::tolower);
    const size_t n = strlen(suf);
    return l.size() >= n && l.compare(l.size() - n, n, suf) == 0;
}

// returns 0 and fills w/h (and rgba when non-null, w*h*4 bytes)
int decode_image(const char* path, int* w, int* h, unsigned char* rgba, size_t capacity)
{
    std::vector<unsigned char> d;
    if (!read_file(path, d)) return io_fail(std::string("The file ") + path + " does not exist.");
    if (ends_with(path, ".tga")) {
        if (d.size() < 18) return io_fail(std::string("Unknown error reading TGA data for ") + path + ".");
        const int idLen = d[0], mapType = d[1], imgType = d[2];
        const int width = d[12] | (d[13] << 8), height = d[14] | (d[15] << 8), depthBits = d[16];
        if (!((imgType == 2 || imgType == 3) && mapType == 0))
            return io_fail(mapType == 1 ? std::string("Does not support TGA color map for ") + path
                                        : std::string("Only supports RGB(A) type TGA textures - does not support type '") + std::to_string(imgType) + "' for " + path);
        const int depth = depthBits / 8;
        const size_t off = 18 + (size_t)idLen;
        if (depth < 1 || d.size() < off + (size_t)width * height * depth) return io_fail(std::string("Unknown error reading TGA data for ") + path + ".");
        *w = width; *h = height;
        if (!rgba) return 0;
        if (capacity < (size_t)width * height * 4) return io_fail("opr_image_read: destination too small");
        std::vector<unsigned char> img(d.begin() + off, d.begin() + off + (size_t)width * height * depth);
        if (depth == 3 || depth == 4)                                          // libtga __TGAbgr2rgb under TGA_RGB
            for (size_t i = 0; i + 2 < img.size(); i += depth) std::swap(img[i], img[i + 2]);
        const size_t last = img.size() - 1;
        for (int y = 0; y < height; y++)
            for (int x = 0; x < width; x++) {                                  // Image.cpp:100-110, no flip, alpha = 1
                const size_t s = ((size_t)y * width + x) * depth, o = ((size_t)y * width + x) * 4;
                rgba[o + 0] = img[std::min(s + 0, last)]; rgba[o + 1] = img[std::min(s + 1, last)]; rgba[o + 2] = img[std::min(s + 2, last)];
                rgba[o + 3] = 1;
            }
        return 0;
    }
    if (ends_with(path, ".bmp")) {
        if (d.size() < 54 || d[0] != 'B' || d[1] != 'M') return io_fail(std::string("Unable to load image ") + path + ".");
        BmpHeader hd; memcpy(&hd, d.data() + 2, sizeof hd);
        const int width = hd.mWidth, height = std::abs(hd.mHeight), bpp = hd.mBitsPerPixel / 8;
        if (hd.mCompression != 0 || (bpp != 3 && bpp != 4) || width <= 0 || height <= 0) return io_fail(std::string("Unable to load image ") + path + ".");
        const size_t stride = ((size_t)width * bpp + 3) & ~(size_t)3;
        if (d.size() < hd.mDataOffset + stride * height) return io_fail(std::string("Unable to load image ") + path + ".");
        *w = width; *h = height;
        if (!rgba) return 0;
        if (capacity < (size_t)width * height * 4) return io_fail("opr_image_read: destination too small");
        for (int y = 0; y < height; y++) {
            const int srcRow = hd.mHeight > 0 ? height - 1 - y : y;           // QImage presents rows top-down
            const unsigned char* row = d.data() + hd.mDataOffset + stride * srcRow;
            for (int x = 0; x < width; x++) {
                unsigned char* o = rgba + ((size_t)y * width + x) * 4;
                o[0] = row[x * bpp + 2]; o[1] = row[x * bpp + 1]; o[2] = row[x * bpp + 0]; o[3] = bpp == 4 ? row[x * bpp + 3] : 255;
            }
        }
        return 0;
    }
    return io_fail(std::string("Unable to load image ") + path + " (supported: .tga, .bmp).");
}

}  // namespace

extern "C" {

const char* opr_io_last_error(void) { return g_ioError.c_str(); }

// the fragment shader of RenderWidget.cpp:187-199 and the per-pixel loop of saveImageAsBMP (:297-315): sum buffer -> 8-bit BGR
int opr_display_bgr8(const float* sumRGB, uint32_t width, uint32_t height, float previewedIterations, float gamma, uint8_t* outBGR)
{
    if (!sumRGB || !outBGR || previewedIterations <= 0.f || gamma <= 0.f) return io_fail("opr_display_bgr8: bad argument");
    const float invGamma = 1.f / gamma;
    const float invIterations = 1.f / previewedIterations;
    for (size_t i = 0; i < (size_t)width * height; i++) {
        const float* c = sumRGB + 3 * i;
        outBGR[3 * i + 0] = to_byte(std::pow(c[2] * invIterations, invGamma) * 255.f);
        outBGR[3 * i + 1] = to_byte(std::pow(c[1] * invIterations, invGamma) * 255.f);
        outBGR[3 * i + 2] = to_byte(std::pow(c[0] * invIterations, invGamma) * 255.f);
    }
    return 0;
}

// RenderWidget::saveImageAsBMP, Gui/gui/RenderWidget.cpp:268-317
int opr_save_bmp(const char* fileName, const float* sumRGB, uint32_t width, uint32_t height, float previewedIterations, float gamma)
{
    std::vector<uint8_t> px((size_t)width * height * 3);
    if (opr_display_bgr8(sumRGB, width, height, previewedIterations, gamma, px.data())) return 1;
    FILE* f = fopen(fileName, "wb");
    if (!f) return io_fail(std::string("cannot open ") + fileName);
    BmpHeader header;
    header.mFileSize = uint32_t(sizeof(BmpHeader) + 2) + width * height * 3;
    header.mReserved01 = 0;
    header.mDataOffset = uint32_t(sizeof(BmpHeader) + 2);
    header.mHeaderSize = 40;
    header.mWidth = (int32_t)width; header.mHeight = (int32_t)height;
    header.mColorPlates = 1; header.mBitsPerPixel = 24; header.mCompression = 0;
    header.mImageSize = width * height * 3;
    header.mHorizRes = 2953; header.mVertRes = 2953; header.mPaletteColors = 0; header.mImportantColors = 0;
    bool ok = fwrite("BM", 1, 2, f) == 2 && fwrite(&header, sizeof header, 1, f) == 1 && fwrite(px.data(), 1, px.size(), f) == px.size();
    fclose(f);
    return ok ? 0 : io_fail(std::string("short write to ") + fileName);
}

int opr_image_info(const char* path, int32_t* width, int32_t* height)
{
    int w = 0, h = 0;
    if (decode_image(path, &w, &h, nullptr, 0)) return 1;
    *width = w; *height = h;
    return 0;
}

int opr_image_read(const char* path, uint8_t* rgba, uint64_t capacityBytes)
{
    int w = 0, h = 0;
    return decode_image(path, &w, &h, rgba, (size_t)capacityBytes);
}

}  // extern "C"
