// generate_matrix <n> <out.mtx> -- reference src/matrix/generate_matrix.cu:36-44
#include "cli_common.h"

int main(int argc, char** argv) {
    if (argc != 3) {
        fprintf(stderr, "Usage: %s <grid_size> <output.mtx>\n", argv[0]);
        return EXIT_FAILURE;
    }
    const int n = atoi(argv[1]);
    if (n <= 0) {
        fprintf(stderr, "grid_size must be a positive integer\n");
        return EXIT_FAILURE;
    }
    return write_matrix_market_stencil5(n, argv[2]) == 0 ? EXIT_SUCCESS : EXIT_FAILURE;
}
