// img_decode.h — image decoding for file-backed scene textures (reference: Texture::load, src/texture.rs:232-243:
// `image::open(path).to_rgba32f()`, run on 32 threads by Scene::load_obj, src/scene.rs:239-254).
//
//   PNG  : decoded here (chunk parser + zlib inflate + the five row filters + Adam7), every colour type and bit depth
//   JPEG : nvJPEG, i.e. decoded on the GPU (baseline and progressive); bound at run time like NCCL, so libptcore has no
//          link dependency on it and a box without libnvjpeg only loses JPEG textures
//   EXR  : decoded here: single-part scan-line files with HALF / FLOAT channels, compression NONE / ZIPS / ZIP / PIZ — what the
//          reference's environment map (res/knob/envmap.exr: 4096 x 2048 FLOAT RGBA, PIZ) needs
//   HDR  : Radiance RGBE (.hdr, "#?RADIANCE" / "#?RGBE"): flat and run-length encoded scan-lines, the -Y +X orientation every
//          writer produces (the reference ships one with its Bistro scene, res/bistro/san_giuseppe_bridge_4k.hdr); value =
//          mantissa * 2^(e - 136) as the `image` crate computes it (no + 0.5)
//   PFM  : raw float maps
// Output follows decode_image of the Python mirror: 8-bit sources stay RGBA8 (the kernels apply v / 255, no sRGB
// decode), 16-bit sources become RGBA32F = v / 65535, float sources stay float; grey -> rgb, missing alpha -> 1.
#pragma once
#include <cstdint>
#include <string>
#include <vector>

namespace ptc {

struct DecodedImage {
    uint32_t width = 0, height = 0;
    uint32_t format = 0;                 // PTC_TEXEL_RGBA8 or PTC_TEXEL_RGBA32F
    std::vector<uint8_t> texels;         // width * height * (4 or 16) bytes
};

enum ImageKind { IMG_UNKNOWN = 0, IMG_PNG, IMG_JPEG, IMG_PFM, IMG_EXR, IMG_HDR };
ImageKind sniff_image(const uint8_t* data, size_t n);

// 0 on success, else `err` says why
int decode_png(const uint8_t* data, size_t n, DecodedImage& out, std::string& err);
int decode_pfm(const uint8_t* data, size_t n, DecodedImage& out, std::string& err);
int decode_exr(const uint8_t* data, size_t n, DecodedImage& out, std::string& err);
int decode_hdr(const uint8_t* data, size_t n, DecodedImage& out, std::string& err);

// nvJPEG decoder: one per host thread (own handle, state, stream and device buffer).
class JpegDecoder {
public:
    explicit JpegDecoder(int cuda_device);
    ~JpegDecoder();
    bool ok() const { return ok_; }
    const std::string& why() const { return why_; }
    int decode(const uint8_t* data, size_t n, DecodedImage& out, std::string& err);
private:
    int device_; bool ok_ = false; std::string why_;
    void* handle_ = nullptr; void* state_ = nullptr; void* stream_ = nullptr;
    unsigned char* d_buf_ = nullptr; size_t d_bytes_ = 0;
};

}  // namespace ptc
