#!/bin/sh
# Builds oracle/_ref/libbowref.so: the reference's OWN DBoW2 lines behind Frame::ComputeBoW (src/Frame.cc:395-402) --
# BowVector.cpp and FeatureVector.cpp whole, FORB::distance and the two TemplatedVocabulary::transform members -- taken at
# build time from where they lie under $REF (never copied into this repo: the generated translation unit lives in a
# temporary directory and only the .so is kept), compiled against oracle/shim_bow/bow_shim.h.  TEST INFRASTRUCTURE only.
set -e
REF=${REF:-/root/reference}
HERE=$(cd "$(dirname "$0")" && pwd)
OUT=$HERE/_ref
D=$REF/Thirdparty/DBoW2/DBoW2
T=$D/TemplatedVocabulary.h
sed -n '1127p' "$T" | grep -q 'void TemplatedVocabulary<TDescriptor,F>::transform(' || { echo "TemplatedVocabulary.h:1127 is not transform(features, v, fv, levelsup)"; exit 1; }
sed -n '1129p' "$T" | grep -q 'FeatureVector &fv' || { echo "TemplatedVocabulary.h:1129 is not the FeatureVector overload"; exit 1; }
sed -n '1194p' "$T" | grep -q '^}' || { echo "TemplatedVocabulary.h:1194 is not the end of transform"; exit 1; }
sed -n '1218p' "$T" | grep -q 'transform(const TDescriptor &feature' || { echo "TemplatedVocabulary.h:1218 is not transform(feature, ...)"; exit 1; }
sed -n '1259p' "$T" | grep -q '^}' || { echo "TemplatedVocabulary.h:1259 is not the end of transform(feature, ...)"; exit 1; }
sed -n '81p' "$D/FORB.cpp" | grep -q 'int FORB::distance' || { echo "FORB.cpp:81 is not FORB::distance"; exit 1; }
TMP=$(mktemp -d)
trap 'rm -rf "$TMP"' EXIT
{
  echo '#include "bow_shim.h"'
  echo 'namespace DBoW2 {'
  sed -n '81,101p' "$D/FORB.cpp"
  sed -n '1126,1194p' "$T"
  sed -n '1217,1259p' "$T"
  echo 'template class TemplatedVocabulary<FORB::TDescriptor, FORB>;'
  echo '}'
} > "$TMP/bow_ref_gen.cpp"
mkdir -p "$OUT"
${CXX:-g++} -std=c++11 -O2 -march=x86-64-v2 -ffp-contract=off -fPIC -shared -w -I"$HERE/shim_bow" -I"$D" \
    "$TMP/bow_ref_gen.cpp" "$D/BowVector.cpp" "$D/FeatureVector.cpp" "$HERE/bow_ref_driver.cpp" -o "$OUT/libbowref.so"
echo "$OUT/libbowref.so"
