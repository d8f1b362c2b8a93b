#!/bin/bash
# TEST INFRASTRUCTURE ONLY. Builds the UNMODIFIED reference (MediaArea/RAWcooked) from the
# sources where they lie under /root/reference into oracle/_ref/ (git-ignored, travels to the
# GPU box with the snapshot):
#   oracle/_ref/rawcooked            the reference CLI (used for --check end-to-end tests)
#   oracle/_ref/libref_ffv1dec.so    reference FFV1 decoder + Transform behind oracle/ref_harness.cpp
# Source list = Project/GNU/CLI/Makefile.am:4-64 (the reference's own build needs autotools,
# which this image lacks, so we compile the same files with the same defines by hand).
set -e
REF=${REF:-/root/reference}
HERE=$(cd "$(dirname "$0")" && pwd)
OUT=$HERE/_ref
OBJ=$OUT/obj
if [ ! -d "$REF/Source" ]; then
  echo "build_ref: $REF not present; keeping prebuilt files in $OUT" >&2
  exit 0
fi
mkdir -p "$OBJ"
S=$REF/Source
INC="-I$S -I$S/Lib/ThirdParty/flac/include -I$S/Lib/ThirdParty/flac/src/libFLAC/include -I$S/Lib/ThirdParty/md5 -I$S/Lib/ThirdParty/thread-pool/include -I$S/Lib/ThirdParty/zlib"
CDEF="-DFLAC__NO_DLL -DFLAC__HAS_OGG=0 -DFLAC__NO_ASM -DHAVE_LROUND=1 -DHAVE_STDINT_H=1 -DHAVE_INTTYPES_H=1"
J=${J:-$(nproc)}
n=0
for f in $(grep -oE 'Source/[^ ]+\.c\b' $REF/Project/GNU/CLI/Makefile.am); do
  o=$OBJ/$(echo $f | tr / _).o
  [ -f $o ] || { gcc -O2 -fPIC -w -c $CDEF $INC $REF/$f -o $o & n=$((n+1)); }
  [ $n -ge $J ] && { wait; n=0; }
done
for f in $(grep -oE 'Source/[^ ]+\.cpp\b' $REF/Project/GNU/CLI/Makefile.am); do
  o=$OBJ/$(echo $f | tr / _).o
  [ -f $o ] || { g++ -std=c++11 -O2 -fPIC -w -pthread -c $INC $REF/$f -o $o & n=$((n+1)); }
  [ $n -ge $J ] && { wait; n=0; }
done
wait
g++ -pthread -o $OUT/rawcooked $OBJ/*.o
# decoder harness: every Lib object + CLI/Input (defines fileinput_issue::ErrorTexts needed by Errors.cpp)
g++ -std=c++11 -O2 -fPIC -w -pthread -c $INC $HERE/ref_harness.cpp -o $OUT/ref_harness.o
g++ -shared -pthread -o $OUT/libref_ffv1dec.so $OUT/ref_harness.o $OBJ/Source_Lib_*.o $OBJ/Source_CLI_Input.cpp.o
echo "build_ref: ok -> $OUT"
