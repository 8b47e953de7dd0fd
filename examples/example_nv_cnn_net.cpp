// The reference's user program (examples/cuda/example_nv_cnn_net.cpp:21-71) against this
// framework: load a converted .anakin.bin, optimise, init a Net<NV, P>, fill the input,
// run prediction() and read the softmax output.  Build:
//   g++ -std=c++17 -I/usr/local/cuda/include examples/example_nv_cnn_net.cpp \
//       -Lanakin_b200/lib -lanakin_b200 -lb200saber -L/usr/local/cuda/lib64 -lcudart \
//       -Wl,-rpath,$PWD/anakin_b200/lib -o example_nv_cnn_net
//   ./example_nv_cnn_net model.anakin.bin [batch] [int8|fp32]
#include <algorithm>
#include <cstdio>
#include <random>
#include <string>
#include <vector>

#include "../anakin_b200/csrc/framework/net.h"

using namespace anakin;
using saber::NV;

template <Precision P>
int run(const std::string& path, int batch) {
    graph::Graph<NV, P> graph;
    Status st = graph.load(path);
    if (!st) { fprintf(stderr, "load failed: %s\n", st.info()); return 1; }
    graph.ResetBatchSize("input_0", batch);
    st = graph.Optimize();
    if (!st) { fprintf(stderr, "Optimize failed: %s\n", st.info()); return 1; }

    Net<NV, P> net;
    st = net.init(graph);
    if (!st) { fprintf(stderr, "Net::init failed: %s\n", st.info()); return 1; }

    auto* d_in = net.get_in("input_0");
    saber::Tensor<saber::NVHX86> h_in(d_in->valid_shape(), saber::AK_FLOAT);
    std::mt19937 rng(42);
    std::uniform_real_distribution<float> dist(-1.f, 1.f);
    float* p = static_cast<float*>(h_in.mutable_data());
    for (long long i = 0; i < h_in.valid_size(); ++i) p[i] = dist(rng);
    d_in->copy_from(h_in, net.stream());

    for (int i = 0; i < 10; ++i) net.prediction();   // warm-up; the 2nd call captures the CUDA graph
    net.sync();
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    const int epoch = 1000;                           // reference benchmark protocol (net_exec_test.cpp:183-231)
    cudaEventRecord(e0, net.stream());
    for (int i = 0; i < epoch; ++i) net.prediction();
    cudaEventRecord(e1, net.stream());
    net.sync();
    float ms = 0.f;
    cudaEventElapsedTime(&ms, e0, e1);
    printf("aveage time %f ms (batch %d, %d ops/prediction)\n", ms / epoch, batch, (int)net.launched_op_count());

    auto* d_out = net.get_out_list()[0];
    saber::Tensor<saber::NVHX86> h_out(d_out->valid_shape(), saber::AK_FLOAT);
    std::vector<float> host(d_out->storage_bytes() / sizeof(float));
    cudaMemcpy(host.data(), d_out->data(), d_out->storage_bytes(), cudaMemcpyDeviceToHost);
    const int classes = d_out->channel(), pitch = d_out->channel_stored();
    for (int n = 0; n < batch; ++n) {
        const float* row = host.data() + static_cast<size_t>(n) * pitch;
        int best = static_cast<int>(std::max_element(row, row + classes) - row);
        printf("image %d: top-1 class %d (p = %f)\n", n, best, row[best]);
    }
    return 0;
}

int main(int argc, char** argv) {
    if (argc < 2) { fprintf(stderr, "usage: %s model.anakin.bin [batch] [int8|fp32|fp16]\n", argv[0]); return 2; }
    const int batch = argc > 2 ? atoi(argv[2]) : 1;
    const std::string prec = argc > 3 ? argv[3] : "fp32";
    if (prec == "int8") return run<Precision::INT8>(argv[1], batch);
    if (prec == "fp16") return run<Precision::FP16>(argv[1], batch);
    return run<Precision::FP32>(argv[1], batch);
}
