#include "PngReader.h"
#include <zlib.h>
#include <cstdio>
#include <cstdlib>
#include <cstring>

static uint32_t be32(const uint8_t* p) { return ((uint32_t)p[0] << 24) | ((uint32_t)p[1] << 16) | ((uint32_t)p[2] << 8) | p[3]; }

bool readPng(const std::string& path, PngImage& out, std::string* err)
{
    auto fail = [&](const char* m) { if (err) *err = path + ": " + m; return false; };
    FILE* f = std::fopen(path.c_str(), "rb");
    if (!f) return fail("cannot open");
    std::vector<uint8_t> buf;
    uint8_t tmp[1 << 16];
    size_t n;
    while ((n = std::fread(tmp, 1, sizeof(tmp), f)) > 0) buf.insert(buf.end(), tmp, tmp + n);
    std::fclose(f);
    static const uint8_t sig[8] = { 0x89, 'P', 'N', 'G', 0x0d, 0x0a, 0x1a, 0x0a };
    if (buf.size() < 33 || std::memcmp(buf.data(), sig, 8) != 0) return fail("not a PNG");
    size_t pos = 8;
    unsigned w = 0, h = 0, depth = 0, ctype = 0, interlace = 0;
    std::vector<uint8_t> idat;
    while (pos + 12 <= buf.size()) {
        const uint32_t len = be32(&buf[pos]);
        const char* type = (const char*)&buf[pos + 4];
        if (pos + 12 + len > buf.size()) return fail("truncated chunk");
        const uint8_t* data = &buf[pos + 8];
        if (!std::memcmp(type, "IHDR", 4)) {
            if (len < 13) return fail("bad IHDR");
            w = be32(data); h = be32(data + 4); depth = data[8]; ctype = data[9]; interlace = data[12];
        } else if (!std::memcmp(type, "IDAT", 4)) {
            idat.insert(idat.end(), data, data + len);
        } else if (!std::memcmp(type, "IEND", 4)) {
            break;
        }
        pos += 12 + len;
    }
    if (!w || !h) return fail("no IHDR");
    if (interlace) return fail("interlaced PNG not supported");
    if (depth != 8 && depth != 16) return fail("only 8/16-bit samples supported");
    unsigned ch;
    switch (ctype) { case 0: ch = 1; break; case 2: ch = 3; break; case 4: ch = 2; break; case 6: ch = 4; break; default: return fail("palette PNG not supported"); }
    const size_t bpp = (size_t)ch * depth / 8, stride = (size_t)w * bpp;
    std::vector<uint8_t> raw((stride + 1) * h);
    uLongf rawlen = raw.size();
    if (uncompress(raw.data(), &rawlen, idat.data(), idat.size()) != Z_OK || rawlen != raw.size()) return fail("zlib inflate failed");
    // reverse the scanline filters in place
    std::vector<uint8_t> prev(stride, 0);
    out.width = w; out.height = h; out.channels = ch; out.bit_depth = depth;
    out.samples.assign((size_t)w * h * ch, 0);
    for (unsigned y = 0; y < h; ++y) {
        uint8_t* line = &raw[(stride + 1) * y];
        const uint8_t ft = line[0];
        uint8_t* cur = line + 1;
        for (size_t i = 0; i < stride; ++i) {
            const int a = i >= bpp ? cur[i - bpp] : 0, b = prev[i], c = i >= bpp ? prev[i - bpp] : 0;
            int add = 0;
            switch (ft) {
            case 0: break;
            case 1: add = a; break;
            case 2: add = b; break;
            case 3: add = (a + b) >> 1; break;
            case 4: { const int p = a + b - c, pa = std::abs(p - a), pb = std::abs(p - b), pc = std::abs(p - c);
                      add = (pa <= pb && pa <= pc) ? a : (pb <= pc ? b : c); break; }
            default: return fail("bad filter type");
            }
            cur[i] = (uint8_t)(cur[i] + add);
        }
        std::memcpy(prev.data(), cur, stride);
        uint16_t* dst = &out.samples[(size_t)y * w * ch];
        if (depth == 16) for (size_t i = 0; i < (size_t)w * ch; ++i) dst[i] = (uint16_t)((cur[2 * i] << 8) | cur[2 * i + 1]);
        else for (size_t i = 0; i < (size_t)w * ch; ++i) dst[i] = cur[i];
    }
    return true;
}
