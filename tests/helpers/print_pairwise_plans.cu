// Host-only helper of tests/test_pairwise_plan.py: prints the leaf plan svm_make_plan builds for every feature count
// given on the command line ("d leaves lo n merges lo n merges ...").  Compiled with nvcc, runs without a GPU.
#include <cstdio>
#include <cstdlib>
#include "../../osml_b200/csrc/svm_decision.cuh"

int main(int argc, char** argv) {
    for (int a = 1; a < argc; ++a) {
        const int d = std::atoi(argv[a]);
        osml::SvmPlan plan;
        if (!osml::svm_make_plan(d, plan)) { std::printf("%d -1\n", d); continue; }
        std::printf("%d %d", d, plan.leaves);
        for (int l = 0; l < plan.leaves; ++l) std::printf(" %d %d %d", plan.leaf[l].lo, plan.leaf[l].n, plan.leaf[l].merges);
        std::printf("\n");
    }
    return 0;
}
