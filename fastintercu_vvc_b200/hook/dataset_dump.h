// dataset_dump.h -- training-data dump hook: writes what the reference's loaders read
// (mlt-cnn-python/codes/data/mlt_ctu_or_pq_dataset.py:13-15,46-69):
//
//   <root>/<sequence>/org/<baseQP>_<POC>_<X>_<Y>.png    16-bit grey PNG, cuw x cuh, the (uint16) cast of the org Pel block
//   <root>/<sequence>/pred/<baseQP>_<POC>_<X>_<Y>.png   the same of the prediction block
//   <csv>: one row per CU   sequence,baseQP,poc,x,y,label,cuQP        (label = PartSplit actually chosen by the RDO)
//
// The reference repository ships only the loader; the dumping side inside its private VTM build was never published.
// This is the host-side (C++14, no dependencies: stored-deflate PNG writer) counterpart, to be called from
// EncCu::xCompressCU next to the predictor with the blocks the hook sees (EncCu.cpp:810-830) and, after the CU's RDO,
// the split that won.  It lets the GPU path be validated and re-trained on real content.
#pragma once
#include <cstdint>
#include <cstdio>
#include <string>

namespace mlt_hook {

// 16-bit greyscale PNG (big-endian samples, filter 0, stored deflate blocks).  Returns false on I/O failure.
bool writePng16(const std::string &path, const uint16_t *pix, int width, int height);

class DatasetDump {
public:
    // creates <root>/<sequence>/{org,pred} and opens <csvPath> for appending
    DatasetDump(const std::string &root, const std::string &sequence, int baseQP, const std::string &csvPath);
    ~DatasetDump();
    bool ok() const { return m_csv != nullptr; }
    // org / pred: top-left Pel of the cuw x cuh block, strides in Pel; label: chosen PartSplit (0 NS, 1 QT, 2 BT_H, 3 BT_V, ...)
    bool dump(const int16_t *org, int orgStride, const int16_t *pred, int predStride, int cuw, int cuh, int poc, int x, int y,
              int label, int cuQP);

private:
    std::string m_root, m_seq;
    int m_baseQP;
    FILE *m_csv = nullptr;
};

} // namespace mlt_hook
