// b200enc: the executable RAWcooked launches instead of ffmpeg (`rawcooked --bin-name b200enc ...`,
// /root/reference/Source/CLI/Global.cpp:543-550). Everything happens in libb200enc.so.
#include "../include/b200enc.h"
int main(int argc, char** argv) { return b200enc_main(argc, argv); }
