// nnue_inference_b200 -- command-line twin of the reference's engine/nnue_inference.cpp:11-65 on top of the C ABI of
// libnnue_b200.so (include/nnue_b200.h): same arguments, same CSV line ("logit_0,...,logit_{C-1},density" at
// std::fixed / setprecision(10)), same exit codes, so that a maintainer can diff the two programs' output directly.
//
//   nnue_inference_b200 <model.nnue> <image.bin> <H> <W>
//
// The image file holds H*W*3 float32 values, read as the engine reads them (HWC view of whatever the caller dumped).
#include <cstdio>
#include <cstdlib>
#include <fstream>
#include <iomanip>
#include <iostream>
#include <string>
#include <vector>

#include "nnue_b200.h"

int main(int argc, char *argv[]) {
    if (argc < 5) {
        std::cerr << "Usage: " << argv[0] << " <model.nnue> <image.bin> <H> <W>" << std::endl;
        return 1;
    }
    const std::string model_path = argv[1], image_path = argv[2];
    const int H = std::atoi(argv[3]), W = std::atoi(argv[4]);
    if (H < 1 || W < 1) {
        std::cerr << "Bad image size" << std::endl;
        return 1;
    }
    const size_t elem_count = static_cast<size_t>(H) * W * 3;
    std::vector<float> image(elem_count);
    std::ifstream img_file(image_path, std::ios::binary);
    if (!img_file.is_open()) {
        std::cerr << "Cannot open image file: " << image_path << std::endl;
        return 1;
    }
    img_file.read(reinterpret_cast<char *>(image.data()), elem_count * sizeof(float));
    if (!img_file) {
        std::cerr << "Failed to read image data" << std::endl;
        return 1;
    }
    nnue_qmodel *model = nullptr;
    if (nnue_q_load(model_path.c_str(), &model) != NNUE_OK) {
        std::cerr << "Failed to load model" << std::endl;
        return 1;
    }
    int32_t dims[8];
    float thr = 0.0f;
    nnue_q_dims(model, dims, &thr);
    std::vector<float> logits(dims[4]);
    float density = 0.0f;
    const int rc = nnue_q_infer_host(model, image.data(), 1, H, W, 0, logits.data(), &density);
    if (rc != NNUE_OK) {
        std::cerr << "Inference failed: " << nnue_error_string(rc) << " " << nnue_last_cuda_error() << std::endl;
        nnue_q_free(model);
        return 1;
    }
    std::cout << std::fixed << std::setprecision(10);
    for (size_t i = 0; i < logits.size(); ++i) std::cout << logits[i] << ",";
    std::cout << density << std::endl;
    nnue_q_free(model);
    return 0;
}
