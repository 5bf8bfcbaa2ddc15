// Host mirror of the Nori educational ray tracer's interfaces (after Nori, Copyright (c) 2015 by Wenzel Jakob);
// re-implemented here without third-party code so that plugins register and parse unchanged -- see DESIGN.md section 1.
// block.cpp -- ImageBlock / BlockGenerator / Bitmap (ref: src/block.cpp:15-152, src/bitmap.cpp:69-122).
#include <fstream>
#include "nori/block.h"

NORI_NAMESPACE_BEGIN

ImageBlock::ImageBlock(const Vector2i &size, const ReconstructionFilter *filter) : m_offset(0, 0), m_size(size) {
    if (filter) {
        /* Tabulate the image reconstruction filter (ref: src/block.cpp:17-27); eval() is the HOST plugin's, so any
           ReconstructionFilter plugin works unchanged -- the table is what crosses the C-ABI (nb_set_filter). */
        m_filterRadius = filter->getRadius();
        m_borderSize = (int) std::ceil(m_filterRadius - 0.5f);
        m_filter = new float[NORI_FILTER_RESOLUTION + 1];
        for (int i = 0; i < NORI_FILTER_RESOLUTION; ++i) {
            float pos = (m_filterRadius * i) / NORI_FILTER_RESOLUTION;
            m_filter[i] = filter->eval(pos);
        }
        m_filter[NORI_FILTER_RESOLUTION] = 0.0f;
        m_lookupFactor = NORI_FILTER_RESOLUTION / m_filterRadius;
    }
    m_data.assign((size_t) rows() * cols() * 4, 0.0f);
}

ImageBlock::~ImageBlock() { delete[] m_filter; }

Bitmap *ImageBlock::toBitmap() const {
    Bitmap *result = new Bitmap(m_size);
    const int c = cols();
    for (int y = 0; y < m_size.y(); ++y) for (int x = 0; x < m_size.x(); ++x) {
        const float *p = &m_data[4 * ((size_t) (y + m_borderSize) * c + (x + m_borderSize))];
        Color3f v = Color4f(p[0], p[1], p[2], p[3]).divideByFilterWeight();
        float *o = result->data() + 3 * ((size_t) y * m_size.x() + x);
        o[0] = v.r(); o[1] = v.g(); o[2] = v.b();
    }
    return result;
}

std::string ImageBlock::toString() const { return format("ImageBlock[offset=%s, size=%s]]", m_offset.toString(), m_size.toString()); }

BlockGenerator::BlockGenerator(const Vector2i &size, int blockSize) : m_size(size), m_blockSize(blockSize) {
    m_numBlocks = Vector2i((int) std::ceil(size.x() / (float) blockSize), (int) std::ceil(size.y() / (float) blockSize));
    m_blocksLeft = m_numBlocks.x() * m_numBlocks.y();
    m_direction = ERight;
    m_block = Point2i(m_numBlocks.x() / 2, m_numBlocks.y() / 2);
    m_stepsLeft = 1;
    m_numSteps = 1;
}

bool BlockGenerator::next(ImageBlock &block) {
    if (m_blocksLeft == 0) return false;
    Point2i pos(m_block.x() * m_blockSize, m_block.y() * m_blockSize);
    block.setOffset(pos);
    block.setSize(Vector2i(std::min(m_size.x() - pos.x(), m_blockSize), std::min(m_size.y() - pos.y(), m_blockSize)));
    if (--m_blocksLeft == 0) return true;
    do {
        switch (m_direction) {
            case ERight: ++m_block.x(); break;
            case EDown: ++m_block.y(); break;
            case ELeft: --m_block.x(); break;
            case EUp: --m_block.y(); break;
        }
        if (--m_stepsLeft == 0) {
            m_direction = (m_direction + 1) % 4;
            if (m_direction == ELeft || m_direction == ERight) ++m_numSteps;
            m_stepsLeft = m_numSteps;
        }
    } while (m_block.x() < 0 || m_block.y() < 0 || m_block.x() >= m_numBlocks.x() || m_block.y() >= m_numBlocks.y());
    return true;
}

// ---------------------------------------------------------------- Bitmap writers
namespace {
template <typename T> void put(std::string &o, T v) { o.append(reinterpret_cast<const char *>(&v), sizeof(T)); }
void putStr(std::string &o, const char *s) { o.append(s, std::strlen(s) + 1); }
void exrAttr(std::string &o, const char *name, const char *type, const std::string &payload) {
    putStr(o, name); putStr(o, type); put<int32_t>(o, (int32_t) payload.size()); o += payload;
}
uint32_t crc32(const unsigned char *d, size_t n, uint32_t crc = 0) {
    static uint32_t table[256]; static bool init = false;
    if (!init) { for (uint32_t i = 0; i < 256; ++i) { uint32_t c = i; for (int k = 0; k < 8; ++k) c = (c & 1) ? 0xedb88320u ^ (c >> 1) : c >> 1; table[i] = c; } init = true; }
    crc = ~crc;
    for (size_t i = 0; i < n; ++i) crc = table[(crc ^ d[i]) & 0xff] ^ (crc >> 8);
    return ~crc;
}
void pngChunk(std::ofstream &os, const char *type, const std::string &data) {
    unsigned char len[4] = { (unsigned char) (data.size() >> 24), (unsigned char) (data.size() >> 16), (unsigned char) (data.size() >> 8), (unsigned char) data.size() };
    os.write((const char *) len, 4);
    std::string td(type, 4); td += data;
    os.write(td.data(), (std::streamsize) td.size());
    uint32_t c = crc32((const unsigned char *) td.data(), td.size());
    unsigned char cb[4] = { (unsigned char) (c >> 24), (unsigned char) (c >> 16), (unsigned char) (c >> 8), (unsigned char) c };
    os.write((const char *) cb, 4);
}
}  // namespace

/// Uncompressed fp32 scanline OpenEXR, channels B G R (the reference writes EXR through OpenEXR, ref: src/bitmap.cpp:69-91)
void Bitmap::saveEXR(const std::string &filename) const {
    std::string path = filename + ".exr";
    const int w = cols(), h = rows();
    std::string hdr;
    put<uint32_t>(hdr, 20000630u); put<uint32_t>(hdr, 2u);
    std::string ch;
    for (const char *c : { "B", "G", "R" }) { putStr(ch, c); put<int32_t>(ch, 2 /*FLOAT*/); put<uint8_t>(ch, 0); ch.append(3, '\0'); put<int32_t>(ch, 1); put<int32_t>(ch, 1); }
    ch.push_back('\0');
    exrAttr(hdr, "channels", "chlist", ch);
    exrAttr(hdr, "compression", "compression", std::string(1, '\0'));
    std::string box; put<int32_t>(box, 0); put<int32_t>(box, 0); put<int32_t>(box, w - 1); put<int32_t>(box, h - 1);
    exrAttr(hdr, "dataWindow", "box2i", box);
    exrAttr(hdr, "displayWindow", "box2i", box);
    exrAttr(hdr, "lineOrder", "lineOrder", std::string(1, '\0'));
    std::string f1; put<float>(f1, 1.0f); exrAttr(hdr, "pixelAspectRatio", "float", f1);
    std::string v2; put<float>(v2, 0.0f); put<float>(v2, 0.0f); exrAttr(hdr, "screenWindowCenter", "v2f", v2);
    exrAttr(hdr, "screenWindowWidth", "float", f1);
    hdr.push_back('\0');
    std::ofstream os(path, std::ios::binary);
    if (!os) throw NoriException("Unable to write \"%s\"", path);
    os.write(hdr.data(), (std::streamsize) hdr.size());
    const uint64_t lineBytes = 8 + (uint64_t) w * 4 * 3;
    uint64_t off = hdr.size() + (uint64_t) h * 8;
    for (int y = 0; y < h; ++y) { os.write((const char *) &off, 8); off += lineBytes; }
    std::vector<float> line((size_t) w * 3);
    for (int y = 0; y < h; ++y) {
        int32_t yy = y, sz = w * 4 * 3;
        os.write((const char *) &yy, 4); os.write((const char *) &sz, 4);
        for (int c = 0; c < 3; ++c) for (int x = 0; x < w; ++x) line[(size_t) c * w + x] = m_px[3 * ((size_t) y * w + x) + (2 - c)];
        os.write((const char *) line.data(), sz);
    }
}

/// The 8-bit sRGB image of savePNG (ref: src/bitmap.cpp:93-122: toSRGB, x255, clamp), 3 bytes per pixel
void Bitmap::toSRGB8(std::vector<uint8_t> &out) const {
    const int w = cols(), h = rows();
    out.resize((size_t) w * h * 3);
    for (size_t i = 0; i < (size_t) w * h; ++i) {
        const float *p = &m_px[3 * i];
        Color3f t = Color3f(p[0], p[1], p[2]).toSRGB();
        for (int c = 0; c < 3; ++c) out[3 * i + c] = (uint8_t) std::min(255.f, std::max(0.f, 255.f * t.c[c]));   // clamp: ref src/bitmap.cpp:107-109
    }
}

void Bitmap::savePNG(const std::string &filename) const {
    std::vector<uint8_t> rgb8;
    toSRGB8(rgb8);
    savePNG8(filename, cols(), rows(), rgb8.data());
}

/// sRGB 8-bit PNG with stored (uncompressed) deflate blocks (ref: src/bitmap.cpp:93-122 uses stb_image_write).  rgb8 is
/// already tonemapped: by toSRGB8 above, or on the device by nb_last_film_to_srgb8 (the same bytes).
void Bitmap::savePNG8(const std::string &filename, int w, int h, const uint8_t *rgb8) {
    std::string path = filename + ".png";
    std::string raw; raw.reserve((size_t) h * (1 + 3 * w));
    for (int y = 0; y < h; ++y) {
        raw.push_back('\0');
        raw.append((const char *) rgb8 + (size_t) y * w * 3, (size_t) w * 3);
    }
    std::string z; z.push_back((char) 0x78); z.push_back((char) 0x01);
    uint32_t a = 1, b = 0;
    for (unsigned char c : raw) { a = (a + c) % 65521u; b = (b + a) % 65521u; }
    for (size_t p = 0; p < raw.size() || p == 0;) {
        size_t n = std::min<size_t>(65535, raw.size() - p);
        z.push_back((char) (p + n >= raw.size() ? 1 : 0));
        z.push_back((char) (n & 0xff)); z.push_back((char) (n >> 8)); z.push_back((char) (~n & 0xff)); z.push_back((char) ((~n >> 8) & 0xff));
        z.append(raw, p, n);
        p += n;
        if (n == 0) break;
    }
    uint32_t ad = (b << 16) | a;
    z.push_back((char) (ad >> 24)); z.push_back((char) (ad >> 16)); z.push_back((char) (ad >> 8)); z.push_back((char) ad);
    std::ofstream os(path, std::ios::binary);
    if (!os) throw NoriException("Unable to write \"%s\"", path);
    const unsigned char sig[8] = { 0x89, 'P', 'N', 'G', '\r', '\n', 0x1a, '\n' };
    os.write((const char *) sig, 8);
    std::string ihdr;
    for (int v : { w, h }) { ihdr.push_back((char) (v >> 24)); ihdr.push_back((char) (v >> 16)); ihdr.push_back((char) (v >> 8)); ihdr.push_back((char) v); }
    ihdr.push_back(8); ihdr.push_back(2); ihdr.push_back(0); ihdr.push_back(0); ihdr.push_back(0);
    pngChunk(os, "IHDR", ihdr); pngChunk(os, "IDAT", z); pngChunk(os, "IEND", "");
}

NORI_NAMESPACE_END
