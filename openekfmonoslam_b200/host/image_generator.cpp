// image_generator.cpp -- FileSequenceImageGenerator (reference: kalmanFilter/modules/ImageGenerator/
// FileSequenceImageGenerator.cpp:60-98) and the image reader behind it (in place of cv::imread): PNG via zlib, PGM / PPM.
#include "../../include/ImageGenerator.h"

#include <zlib.h>

#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <iostream>
#include <vector>

namespace {

typedef unsigned char u8;

bool read_file(const char* name, std::vector<u8>& out)
{
    FILE* f = std::fopen(name, "rb");
    if (!f) return false;
    std::fseek(f, 0, SEEK_END);
    const long n = std::ftell(f);
    std::fseek(f, 0, SEEK_SET);
    out.resize(n > 0 ? (size_t)n : 0);
    const bool ok = n >= 0 && std::fread(out.data(), 1, out.size(), f) == out.size();
    std::fclose(f);
    return ok;
}

inline unsigned be32(const u8* p) { return ((unsigned)p[0] << 24) | ((unsigned)p[1] << 16) | ((unsigned)p[2] << 8) | p[3]; }

// output: always 3 interleaved bytes per pixel, blue first (cv::imread's default)
void put_bgr(cv::Mat& m, int y, int x, u8 r, u8 g, u8 b)
{
    u8* p = m.ptr<u8>(y) + 3 * x;
    p[0] = b; p[1] = g; p[2] = r;
}

bool decode_png(const std::vector<u8>& file, cv::Mat& out, const char** why)
{
    static const u8 sig[8] = {0x89, 'P', 'N', 'G', 0x0d, 0x0a, 0x1a, 0x0a};
    if (file.size() < 33 || std::memcmp(file.data(), sig, 8) != 0) { *why = "not a PNG file"; return false; }
    unsigned W = 0, H = 0;
    int depth = 0, ctype = -1, interlace = 0;
    std::vector<u8> idat, plte;
    for (size_t pos = 8; pos + 12 <= file.size();) {
        const unsigned len = be32(&file[pos]);
        const u8* type = &file[pos + 4];
        const u8* data = &file[pos + 8];
        if (pos + 12 + (size_t)len > file.size()) { *why = "truncated PNG chunk"; return false; }
        if (!std::memcmp(type, "IHDR", 4) && len >= 13) {
            W = be32(data); H = be32(data + 4);
            depth = data[8]; ctype = data[9]; interlace = data[12];
        } else if (!std::memcmp(type, "PLTE", 4)) {
            plte.assign(data, data + len);
        } else if (!std::memcmp(type, "IDAT", 4)) {
            idat.insert(idat.end(), data, data + len);
        } else if (!std::memcmp(type, "IEND", 4)) {
            break;
        }
        pos += 12 + (size_t)len;
    }
    if (W == 0 || H == 0 || W > 16384 || H > 16384) { *why = "bad PNG dimensions"; return false; }
    if (depth != 8) { *why = "only 8-bit PNG samples are handled"; return false; }
    if (interlace != 0) { *why = "interlaced PNG is not handled"; return false; }
    int ch;
    switch (ctype) {
        case 0: ch = 1; break;   // grey
        case 2: ch = 3; break;   // RGB
        case 3: ch = 1; break;   // palette index
        case 4: ch = 2; break;   // grey + alpha
        case 6: ch = 4; break;   // RGBA
        default: *why = "unknown PNG colour type"; return false;
    }
    if (ctype == 3 && plte.size() < 3) { *why = "palette PNG without PLTE"; return false; }
    const size_t stride = (size_t)W * ch;
    std::vector<u8> raw((stride + 1) * H);
    uLongf rawLen = (uLongf)raw.size();
    if (uncompress(raw.data(), &rawLen, idat.data(), (uLong)idat.size()) != Z_OK || rawLen != raw.size()) {
        *why = "PNG data does not inflate to the image size";
        return false;
    }
    // undo the per-row filters in place (PNG specification, section 9: None, Sub, Up, Average, Paeth)
    std::vector<u8> zero(stride, 0);
    for (unsigned y = 0; y < H; ++y) {
        u8* row = &raw[(stride + 1) * y + 1];
        const u8* up = y ? &raw[(stride + 1) * (y - 1) + 1] : zero.data();
        const int ft = raw[(stride + 1) * y];
        for (size_t i = 0; i < stride; ++i) {
            const int a = i >= (size_t)ch ? row[i - ch] : 0, b = up[i], c = i >= (size_t)ch ? up[i - ch] : 0;
            int pred = 0;
            if (ft == 1) pred = a;
            else if (ft == 2) pred = b;
            else if (ft == 3) pred = (a + b) >> 1;
            else if (ft == 4) {
                const int p = a + b - c, pa = std::abs(p - a), pb = std::abs(p - b), pc = std::abs(p - c);
                pred = (pa <= pb && pa <= pc) ? a : (pb <= pc ? b : c);
            } else if (ft != 0) { *why = "bad PNG filter type"; return false; }
            row[i] = (u8)(row[i] + pred);
        }
    }
    out = cv::Mat((int)H, (int)W, CV_8UC3);
    for (unsigned y = 0; y < H; ++y) {
        const u8* row = &raw[(stride + 1) * y + 1];
        for (unsigned x = 0; x < W; ++x) {
            const u8* p = row + (size_t)x * ch;
            if (ctype == 0 || ctype == 4) put_bgr(out, y, x, p[0], p[0], p[0]);
            else if (ctype == 3) {
                const size_t e = (size_t)p[0] * 3;
                if (e + 2 < plte.size()) put_bgr(out, y, x, plte[e], plte[e + 1], plte[e + 2]);
                else put_bgr(out, y, x, 0, 0, 0);
            } else put_bgr(out, y, x, p[0], p[1], p[2]);
        }
    }
    return true;
}

// binary PGM (P5) / PPM (P6), maxval <= 255
bool decode_pnm(const std::vector<u8>& file, cv::Mat& out, const char** why)
{
    if (file.size() < 7 || file[0] != 'P' || (file[1] != '5' && file[1] != '6')) { *why = "not a binary PGM / PPM file"; return false; }
    const int ch = file[1] == '6' ? 3 : 1;
    size_t pos = 2;
    int vals[3], got = 0;
    while (got < 3 && pos < file.size()) {
        if (file[pos] == '#') { while (pos < file.size() && file[pos] != '\n') ++pos; continue; }
        if (file[pos] <= ' ') { ++pos; continue; }
        int v = 0;
        while (pos < file.size() && file[pos] >= '0' && file[pos] <= '9') v = v * 10 + (file[pos++] - '0');
        vals[got++] = v;
    }
    ++pos;   // the single whitespace byte behind maxval
    if (got < 3 || vals[0] <= 0 || vals[1] <= 0 || vals[2] <= 0 || vals[2] > 255) { *why = "bad PNM header"; return false; }
    const size_t need = (size_t)vals[0] * vals[1] * ch;
    if (pos + need > file.size()) { *why = "truncated PNM data"; return false; }
    out = cv::Mat(vals[1], vals[0], CV_8UC3);
    const u8* p = &file[pos];
    for (int y = 0; y < vals[1]; ++y)
        for (int x = 0; x < vals[0]; ++x, p += ch) put_bgr(out, y, x, p[0], p[ch == 3 ? 1 : 0], p[ch == 3 ? 2 : 0]);
    return true;
}

}  // namespace

bool ekfbReadImage(const char* fileName, cv::Mat& bgr)
{
    bgr = cv::Mat();
    std::vector<u8> file;
    const char* why = "cannot open the file";
    bool ok = fileName != nullptr && read_file(fileName, file);
    if (ok) ok = (file.size() > 1 && file[0] == 'P') ? decode_pnm(file, bgr, &why) : decode_png(file, bgr, &why);
    if (!ok) {
        std::cerr << "Unable to read image: " << (fileName ? fileName : "(null)") << " (" << why << ")" << std::endl;
        bgr = cv::Mat();
    }
    return ok;
}

FileSequenceImageGenerator::FileSequenceImageGenerator() : _imageBeginIndex(0), _imageEndIndex(-1), _imageActualIndex(0) {}

FileSequenceImageGenerator::FileSequenceImageGenerator(std::string path, std::string filePrefix, std::string fileExtension,
                                                       int imageBeginIndex, int imageEndIndex)
    : _path(path), _filePrefix(filePrefix), _fileExtension(fileExtension), _imageBeginIndex(imageBeginIndex),
      _imageEndIndex(imageEndIndex), _imageActualIndex(imageBeginIndex)
{
}

FileSequenceImageGenerator::~FileSequenceImageGenerator() {}

void FileSequenceImageGenerator::init() { _imageActualIndex = _imageBeginIndex; }

// "<path><prefix>%05d.<ext>" for the next index; past the last index -- or when the file cannot be read -- an empty image,
// which ends the caller's loop (kalmanFilter/samples/EKF/main.cpp:133)
cv::Mat& FileSequenceImageGenerator::getNextImage()
{
    if (_imageActualIndex > _imageEndIndex) {
        _image = cv::Mat();
        return _image;
    }
    char index[16];
    std::snprintf(index, sizeof(index), "%05d", _imageActualIndex);
    const std::string name = _path + _filePrefix + index + "." + _fileExtension;
    ekfbReadImage(name.c_str(), _image);
    _imageActualIndex++;
    return _image;
}

// flat C hooks (ctypes tests, JNI-style callers): decode a file into a caller buffer of 3 * w * h bytes (BGR); the first call
// with out == NULL returns the size
extern "C" int ekfb_host_read_image(const char* fileName, unsigned char* out, int* width, int* height)
{
    cv::Mat m;
    if (!ekfbReadImage(fileName, m)) return 1;
    if (width) *width = m.cols;
    if (height) *height = m.rows;
    if (out)
        for (int y = 0; y < m.rows; ++y) std::memcpy(out + (size_t)y * m.cols * 3, m.ptr<unsigned char>(y), (size_t)m.cols * 3);
    return 0;
}
