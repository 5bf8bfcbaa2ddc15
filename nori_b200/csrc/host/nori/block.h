// Host mirror of the Nori educational ray tracer's interfaces (after Nori, Copyright (c) 2015 by Wenzel Jakob);
// re-implemented here without third-party code so that plugins register and parse unchanged -- see DESIGN.md section 1.
// block.h -- ImageBlock: weighted RGBA film storage incl. border (ref: include/nori/block.h:31-112, src/block.cpp:15-51).
// On this path the block is FILLED by the GPU (nb_render writes the byte-compatible storage); the host keeps the
// constructor (filter tabulation through the host plugin's eval()), toBitmap and the BlockGenerator for API parity.
#pragma once
#include "plugins.h"

#define NORI_BLOCK_SIZE 32          /* ref: include/nori/block.h:17 */
#define NORI_FILTER_RESOLUTION 32   /* ref: include/nori/rfilter.h:12 */

NORI_NAMESPACE_BEGIN

class Bitmap;

class ImageBlock {
public:
    ImageBlock(const Vector2i &size, const ReconstructionFilter *filter);
    ~ImageBlock();
    void setOffset(const Point2i &offset) { m_offset = offset; }
    const Point2i &getOffset() const { return m_offset; }
    void setSize(const Point2i &size) { m_size = size; }
    const Vector2i &getSize() const { return m_size; }
    int getBorderSize() const { return m_borderSize; }
    int rows() const { return m_size.y() + 2 * m_borderSize; }
    int cols() const { return m_size.x() + 2 * m_borderSize; }
    void clear() { std::fill(m_data.begin(), m_data.end(), 0.0f); }
    Bitmap *toBitmap() const;                                  // ref: src/block.cpp:45-51
    float *data() { return m_data.data(); }                    // rows() x cols() x 4, row-major
    const float *data() const { return m_data.data(); }
    const float *filterTable() const { return m_filter; }      // NORI_FILTER_RESOLUTION + 1 entries (ref: src/block.cpp:21-26)
    float filterRadius() const { return m_filterRadius; }
    std::string toString() const;
protected:
    Point2i m_offset;
    Vector2i m_size;
    int m_borderSize = 0;
    float *m_filter = nullptr;
    float m_filterRadius = 0;
    float m_lookupFactor = 0;
    std::vector<float> m_data;
};

/// Spiral tile scheduler (ref: include/nori/block.h:118-152, src/block.cpp:109-152).  The GPU path shards tiles by
/// tile_id % nGPU instead; kept because the block order is part of the reference's public surface (GUI).
class BlockGenerator {
public:
    BlockGenerator(const Vector2i &size, int blockSize);
    bool next(ImageBlock &block);
    int getBlockCount() const { return m_blocksLeft; }
protected:
    enum EDirection { ERight = 0, EDown, ELeft, EUp };
    Point2i m_block; Vector2i m_numBlocks; Vector2i m_size;
    int m_blockSize, m_numSteps, m_blocksLeft, m_stepsLeft, m_direction;
};

/// RGB bitmap (ref: include/nori/bitmap.h:17-40): rows x cols of Color3f; EXR + PNG writers without OpenEXR / stb.
class Bitmap {
public:
    Bitmap(const Vector2i &size) : m_size(size), m_px((size_t) size.x() * size.y() * 3, 0.0f) { }
    int cols() const { return m_size.x(); }
    int rows() const { return m_size.y(); }
    float *data() { return m_px.data(); }
    const float *data() const { return m_px.data(); }
    void saveEXR(const std::string &filename) const;            // uncompressed fp32 scanline OpenEXR ("<name>.exr")
    void savePNG(const std::string &filename) const;            // sRGB 8-bit ("<name>.png"), ref: src/bitmap.cpp:93-122
    void toSRGB8(std::vector<uint8_t> &out) const;              // the tonemapped bytes savePNG writes (host loop)
    static void savePNG8(const std::string &filename, int w, int h, const uint8_t *rgb8);   // PNG of already tonemapped bytes
private:
    Vector2i m_size;
    std::vector<float> m_px;
};

NORI_NAMESPACE_END
