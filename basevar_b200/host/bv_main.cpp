// bv_main.cpp -- `basevar basetype` with the per-site statistics on the GPU: same options and output files as the
// reference's command (src/main.cpp:40-53, src/basetype_caller.cpp:19-142); BAM/FASTA in, VCF/CVG out.
#include <cstring>
#include <iostream>

#include "bv_pileup.hpp"

int main(int argc, char* argv[]) {
    if (argc < 2 || strcmp(argv[1], "basetype") != 0) {
        std::cerr << "Usage: basevar basetype [options]   (only the basetype command is built here)\n" << bvhost::BaseTypeRunner::usage() << std::endl;
        return 1;
    }
    try {
        bvhost::BaseTypeRunner runner(argc - 1, argv + 1);
        runner.run();
    } catch (const std::exception& e) {
        std::cerr << e.what() << std::endl;
        return 1;
    }
    return 0;
}
