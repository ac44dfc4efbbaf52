// dump_tool.cpp -- test driver of dataset_dump.h: reads "n size" then n records {int32 poc, x, y, label, cuQP; int16 org[size*size];
// int16 pred[size*size]} from a file and dumps them.  usage: dump_tool in.bin root sequence baseQP csv
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "dataset_dump.h"

int main(int argc, char **argv)
{
    if (argc != 6) return 2;
    FILE *f = std::fopen(argv[1], "rb");
    if (!f) return 2;
    int hdr[2];
    if (std::fread(hdr, sizeof(int), 2, f) != 2) return 2;
    const int n = hdr[0], size = hdr[1], stride = size + 24; // strided source buffers, like a picture
    mlt_hook::DatasetDump d(argv[2], argv[3], std::atoi(argv[4]), argv[5]);
    if (!d.ok()) return 3;
    std::vector<int16_t> blk((size_t)size * size), o((size_t)size * stride), p((size_t)size * stride);
    for (int i = 0; i < n; i++) {
        int meta[5];
        if (std::fread(meta, sizeof(int), 5, f) != 5) return 2;
        for (std::vector<int16_t> *dst : {&o, &p}) {
            if (std::fread(blk.data(), sizeof(int16_t), blk.size(), f) != blk.size()) return 2;
            for (int y = 0; y < size; y++)
                for (int x = 0; x < size; x++) (*dst)[(size_t)y * stride + x] = blk[(size_t)y * size + x];
        }
        if (!d.dump(o.data(), stride, p.data(), stride, size, size, meta[0], meta[1], meta[2], meta[3], meta[4])) return 4;
    }
    std::fclose(f);
    std::printf("dumped %d\n", n);
    return 0;
}
