// b200enc: the executable RAWcooked launches instead of ffmpeg (`rawcooked --bin-name b200enc ...`,
// /root/reference/Source/CLI/Global.cpp:543-550). Everything happens in libb200enc.so.
#include <unistd.h>

#include <cstdio>

#include "../include/b200enc.h"
int main(int argc, char** argv) {
    b200enc_set_fast_exit(1);            // the process ends with the job: no freeing of device memory, no unpinning
    const int rc = b200enc_main(argc, argv);
    fflush(stdout);
    fflush(stderr);
    _exit(rc & 0xFF ? rc & 0xFF : (rc ? 1 : 0));
}
