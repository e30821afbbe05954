/* papr_cli.c — `papr [-g] <infile>`: drop-in executable for drmpeg/dtv-utils papr (papr.c:32). */
#include "../../include/papr_b200.h"
int main(int argc, char **argv) { return papr_main(argc, argv); }
