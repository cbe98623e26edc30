// choh_batch — host-side .hoh container writer on top of libhohgpu.so.
//
// The C++ host program BASELINE.json's north_star describes: the container (file header, tile grid, tile-size
// varints, concatenation — choh.cpp:436-506, varint.hpp:29-45) stays on the host in C++, and everything inside a
// tile (LZ finder, colour planes, prediction, predictor search, rANS, tile assembly: encode_tile choh.cpp:104-382)
// is ONE call into the C-ABI for the whole batch: hoh_encode_images_host.
//
//   choh_batch infile.rgb outfile.hoh width height -sN          the reference tool's own command line (choh.cpp:394-433)
//   choh_batch --batch width height -sN outdir in1.rgb [in2.rgb ...]
//                                                               many images of one size in one GPU call -> outdir/<name>.hoh
// Options (anywhere): --decodable   HOH_FIX_ENCODER: the variant hoh_decode_images / dhoh_batch can invert (stock choh's
//                                   output at -s1..4 is not decodable where SURVEY D7 strikes), and an untiled image's
//                                   tile is written (stock choh sizes it but never writes it)
//                     --device N    CUDA device
//                     --strict      fail (exit 4) when a tile is grey / has <= 256 colours: the reference would code it
//                                   in its greyscale / indexed mode, which stays on the host (default: warn on stderr)
// Without flags the bytes are stock `choh`'s, and like choh the total size is printed on stdout, one line per image.
// Exit codes: 0 ok, 1 usage, 2 I/O, 3 no GPU / library error, 4 --strict violation.
#include <cerrno>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "hohgpu.h"

namespace {

void usage() {
    std::fprintf(stderr,
                 "usage: choh_batch infile.rgb outfile.hoh width height -sN\n"
                 "       choh_batch --batch width height -sN outdir in1.rgb [in2.rgb ...]\n"
                 "options: --decodable  --device N  --strict\n");
}

// varint.hpp:29-45: big-endian 7-bit groups, three bytes at most; larger values emit nothing
void put_varint(std::vector<uint8_t>& out, size_t v) {
    if (v < (1u << 7)) {
        out.push_back((uint8_t)v);
    } else if (v < (1u << 14)) {
        out.push_back((uint8_t)(0x80 | (v >> 7)));
        out.push_back((uint8_t)(v & 0x7f));
    } else if (v < (1u << 21)) {
        out.push_back((uint8_t)(0x80 | (v >> 14)));
        out.push_back((uint8_t)(0x80 | ((v >> 7) & 0x7f)));
        out.push_back((uint8_t)(v & 0x7f));
    }
}

bool read_exact(const char* path, uint8_t* dst, size_t want) {
    FILE* f = std::fopen(path, "rb");
    if (!f) {
        std::fprintf(stderr, "choh_batch: cannot open %s: %s\n", path, std::strerror(errno));
        return false;
    }
    const size_t got = std::fread(dst, 1, want, f);
    uint8_t extra;
    const bool longer = got == want && std::fread(&extra, 1, 1, f) == 1;
    std::fclose(f);
    if (got != want || longer) {
        std::fprintf(stderr, "choh_batch: %s does not hold exactly %zu bytes (width*height*3)\n", path, want);
        return false;
    }
    return true;
}

bool write_all(const std::string& path, const std::vector<uint8_t>& data) {
    FILE* f = std::fopen(path.c_str(), "wb");
    if (!f) {
        std::fprintf(stderr, "choh_batch: cannot create %s: %s\n", path.c_str(), std::strerror(errno));
        return false;
    }
    const bool ok = std::fwrite(data.data(), 1, data.size(), f) == data.size();
    return std::fclose(f) == 0 && ok;
}

std::string stem_of(const std::string& path) {
    const size_t slash = path.find_last_of('/');
    std::string name = slash == std::string::npos ? path : path.substr(slash + 1);
    const size_t dot = name.find_last_of('.');
    return dot == std::string::npos || dot == 0 ? name : name.substr(0, dot);
}

}  // namespace

int main(int argc, char** argv) {
    bool batch = false, decodable = false, strict = false;
    int device = 0;
    std::vector<std::string> pos;
    for (int i = 1; i < argc; i++) {
        const std::string a = argv[i];
        if (a == "--batch") batch = true;
        else if (a == "--decodable") decodable = true;
        else if (a == "--strict") strict = true;
        else if (a == "--device" && i + 1 < argc) device = std::atoi(argv[++i]);
        else pos.push_back(a);
    }
    std::vector<std::string> inputs, outputs;
    int width = 0, height = 0;
    std::string speed;
    if (!batch) {
        if (pos.size() < 4) {
            std::printf("not enough arguments\n");  // choh.cpp:396
            usage();
            return 1;
        }
        inputs.push_back(pos[0]);
        outputs.push_back(pos[1]);
        width = std::atoi(pos[2].c_str());
        height = std::atoi(pos[3].c_str());
        speed = pos.size() > 4 ? pos[4] : "-s1";  // choh.cpp:410: default cruncher mode 1
    } else {
        if (pos.size() < 5) {
            usage();
            return 1;
        }
        width = std::atoi(pos[0].c_str());
        height = std::atoi(pos[1].c_str());
        speed = pos[2];
        for (size_t i = 4; i < pos.size(); i++) {
            inputs.push_back(pos[i]);
            outputs.push_back(pos[3] + "/" + stem_of(pos[i]) + ".hoh");
        }
    }
    if (width <= 0 || height <= 0) {
        std::printf("invalid width or height\n");  // choh.cpp:404
        usage();
        return 1;
    }
    if (speed.size() != 3 || speed[0] != '-' || speed[1] != 's' || speed[2] < '0' || speed[2] > '4') {
        std::printf("invalid speed setting\n");  // choh.cpp:427
        usage();
        return 1;
    }
    const int mode = speed[2] - '0';

    hoh_ctx* ctx = nullptr;
    int st = hoh_ctx_create(device, nullptr, &ctx);
    if (st != HOH_OK) {
        std::fprintf(stderr, "choh_batch: no usable CUDA device %d: %s (there is no CPU fallback)\n", device, hoh_strerror(st));
        return 3;
    }
    hoh_tile_geometry g;
    hoh_tile_geometry_for((uint32_t)width, (uint32_t)height, &g);
    const size_t n = inputs.size(), raw1 = (size_t)width * height * 3, n_tiles = n * g.tiles_per_image;
    const size_t packed_cap = n * raw1 + n * raw1 / 2 + 8192 * n_tiles;
    uint8_t *rgb = nullptr, *packed = nullptr;
    if (hoh_host_alloc(ctx, n * raw1, (void**)&rgb) != HOH_OK || hoh_host_alloc(ctx, packed_cap, (void**)&packed) != HOH_OK) {
        std::fprintf(stderr, "choh_batch: cannot pin host memory: %s\n", hoh_last_cuda_error(ctx));
        return 3;
    }
    for (size_t i = 0; i < n; i++)
        if (!read_exact(inputs[i].c_str(), rgb + i * raw1, raw1)) return 2;
    std::vector<uint64_t> off(n_tiles + 1);
    std::vector<hoh_tile_result> rec(n_tiles);
    st = hoh_encode_images_host(ctx, rgb, n, (uint32_t)width, (uint32_t)height, mode, decodable ? HOH_FIX_ENCODER : 0u, packed,
                                packed_cap, off.data(), rec.data());
    if (st != HOH_OK) {
        std::fprintf(stderr, "choh_batch: hoh_encode_images_host: %s (%s)\n", hoh_strerror(st), hoh_last_cuda_error(ctx));
        return 3;
    }
    int rc = 0;
    const bool tiled = g.tiles_per_image > 1 || ((width >= 512 || height >= 512) && width >= 256 && height >= 256);
    for (size_t i = 0; i < n && rc == 0; i++) {
        const size_t t0 = i * g.tiles_per_image;
        for (size_t t = t0; t < t0 + g.tiles_per_image; t++) {
            if (rec[t].status != HOH_S_OK) {
                std::fprintf(stderr, "choh_batch: %s tile %zu failed with stream status %d\n", inputs[i].c_str(), t - t0, rec[t].status);
                rc = 3;
            }
            if (rec[t].flags) {
                std::fprintf(stderr, "choh_batch: %s tile %zu is %s: the reference codes it in %s mode on the host; written in subtract-green mode\n",
                             inputs[i].c_str(), t - t0, (rec[t].flags & HOH_TILE_GREY) ? "grey" : "a <= 256 colour tile",
                             (rec[t].flags & HOH_TILE_GREY) ? "its greyscale" : "(possibly) its indexed");
                if (strict) rc = 4;
            }
        }
        if (rc) break;
        std::vector<uint8_t> file;  // choh.cpp:436-451
        file.reserve(raw1 / 2 + 64);
        file.push_back(153), file.push_back(72), file.push_back(79), file.push_back(72);
        file.push_back(2);  // RGB
        file.push_back(8);  // bits per channel
        put_varint(file, (size_t)width - 1);
        put_varint(file, (size_t)height - 1);
        size_t printed;
        if (tiled) {  // choh.cpp:454-506
            file.push_back((uint8_t)(g.x_tiles - 1));
            file.push_back((uint8_t)(g.y_tiles - 1));
            for (size_t t = t0; t + 1 < t0 + g.tiles_per_image; t++) put_varint(file, (size_t)(off[t + 1] - off[t]));
            file.insert(file.end(), packed + off[t0], packed + off[t0 + g.tiles_per_image]);
            printed = file.size();
        } else {  // choh.cpp:508-519: the tile is encoded, counted and dropped (D1); --decodable keeps it
            printed = file.size() + (size_t)(off[t0 + 1] - off[t0]);
            if (decodable) file.insert(file.end(), packed + off[t0], packed + off[t0 + 1]);
        }
        std::printf("%d\n", (int)printed);  // choh.cpp:521
        if (!write_all(outputs[i], file)) rc = 2;
    }
    hoh_host_free(ctx, rgb);
    hoh_host_free(ctx, packed);
    hoh_ctx_destroy(ctx);
    return rc;
}
