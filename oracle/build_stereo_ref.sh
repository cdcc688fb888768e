#!/bin/sh
# Builds oracle/_ref/libstereoref.so: the reference's OWN lines of Frame::ComputeStereoMatches,
# ORBmatcher::DescriptorDistance, the Frame undistort/grid functions, Frame::GetFeaturesInArea and
# ORBmatcher::SearchByProjection(Frame&, const Frame&) / ComputeThreeMaxima, SearchByProjection(Frame&, vpMapPoints), SearchByBoW(KeyFrame*,
# Frame&), SearchForInitialization, SearchByProjection(Frame&, KeyFrame*, sAlreadyFound, th, ORBdist), MapPoint::PredictScale and Frame::isInFrustum, taken at build time from where they lie under $REF (never copied into this
# repo: the generated translation unit lives in a temporary directory and only the .so is kept), compiled against
# oracle/shim_stereo/stereo_shim.h.  TEST INFRASTRUCTURE only.
set -e
REF=${REF:-/root/reference}
HERE=$(cd "$(dirname "$0")" && pwd)
OUT=$HERE/_ref
F=$REF/src/Frame.cc
M=$REF/src/ORBmatcher.cc
P=$REF/src/MapPoint.cc
# the line ranges are pinned to the reference revision surveyed in SURVEY.md; refuse to build from anything else
sed -n '466p' "$F" | grep -q 'void Frame::ComputeStereoMatches()' || { echo "Frame.cc:466 is not ComputeStereoMatches"; exit 1; }
sed -n '640p' "$F" | grep -q '^}' || { echo "Frame.cc:640 is not the end of ComputeStereoMatches"; exit 1; }
sed -n '1647p' "$M" | grep -q 'int ORBmatcher::DescriptorDistance' || { echo "ORBmatcher.cc:1647 is not DescriptorDistance"; exit 1; }
sed -n '230p' "$F" | grep -q 'void Frame::AssignFeaturesToGrid()' || { echo "Frame.cc:230 is not AssignFeaturesToGrid"; exit 1; }
sed -n '382p' "$F" | grep -q 'bool Frame::PosInGrid' || { echo "Frame.cc:382 is not PosInGrid"; exit 1; }
sed -n '404p' "$F" | grep -q 'void Frame::UndistortKeyPoints()' || { echo "Frame.cc:404 is not UndistortKeyPoints"; exit 1; }
sed -n '436p' "$F" | grep -q 'void Frame::ComputeImageBounds' || { echo "Frame.cc:436 is not ComputeImageBounds"; exit 1; }
sed -n '327p' "$F" | grep -q 'Frame::GetFeaturesInArea' || { echo "Frame.cc:327 is not GetFeaturesInArea"; exit 1; }
sed -n '1328p' "$M" | grep -q 'int ORBmatcher::SearchByProjection(Frame &CurrentFrame, const Frame &LastFrame' || { echo "ORBmatcher.cc:1328 is not SearchByProjection(Frame&, const Frame&)"; exit 1; }
sed -n '1470p' "$M" | grep -q '^}' || { echo "ORBmatcher.cc:1470 is not the end of SearchByProjection"; exit 1; }
sed -n '1601p' "$M" | grep -q 'void ORBmatcher::ComputeThreeMaxima' || { echo "ORBmatcher.cc:1601 is not ComputeThreeMaxima"; exit 1; }
sed -n '45p' "$M" | grep -q 'int ORBmatcher::SearchByProjection(Frame &F, const vector<MapPoint\*> &vpMapPoints' || { echo "ORBmatcher.cc:45 is not SearchByProjection(Frame&, vpMapPoints)"; exit 1; }
sed -n '129p' "$M" | grep -q '^}' || { echo "ORBmatcher.cc:129 is not the end of SearchByProjection(Frame&, vpMapPoints)"; exit 1; }
sed -n '131p' "$M" | grep -q 'float ORBmatcher::RadiusByViewingCos' || { echo "ORBmatcher.cc:131 is not RadiusByViewingCos"; exit 1; }
sed -n '159p' "$M" | grep -q 'int ORBmatcher::SearchByBoW(KeyFrame\* pKF,Frame &F' || { echo "ORBmatcher.cc:159 is not SearchByBoW(KeyFrame*, Frame&)"; exit 1; }
sed -n '288p' "$M" | grep -q '^}' || { echo "ORBmatcher.cc:288 is not the end of SearchByBoW(KeyFrame*, Frame&)"; exit 1; }
sed -n '405p' "$M" | grep -q 'int ORBmatcher::SearchForInitialization' || { echo "ORBmatcher.cc:405 is not SearchForInitialization"; exit 1; }
sed -n '520p' "$M" | grep -q '^}' || { echo "ORBmatcher.cc:520 is not the end of SearchForInitialization"; exit 1; }
sed -n '1472p' "$M" | grep -q 'int ORBmatcher::SearchByProjection(Frame &CurrentFrame, KeyFrame \*pKF' || { echo "ORBmatcher.cc:1472 is not SearchByProjection(Frame&, KeyFrame*, ...)"; exit 1; }
sed -n '1599p' "$M" | grep -q '^}' || { echo "ORBmatcher.cc:1599 is not the end of SearchByProjection(Frame&, KeyFrame*, ...)"; exit 1; }
sed -n '373p' "$P" | grep -q 'float MapPoint::GetMinDistanceInvariance' || { echo "MapPoint.cc:373 is not GetMinDistanceInvariance"; exit 1; }
sed -n '402p' "$P" | grep -q 'int MapPoint::PredictScale(const float &currentDist, Frame\* pF)' || { echo "MapPoint.cc:402 is not PredictScale(dist, Frame*)"; exit 1; }
sed -n '417p' "$P" | grep -q '^}' || { echo "MapPoint.cc:417 is not the end of PredictScale"; exit 1; }
sed -n '269p' "$F" | grep -q 'bool Frame::isInFrustum(MapPoint \*pMP, float viewingCosLimit)' || { echo "Frame.cc:269 is not isInFrustum"; exit 1; }
sed -n '325p' "$F" | grep -q '^}' || { echo "Frame.cc:325 is not the end of isInFrustum"; exit 1; }
TMP=$(mktemp -d)
trap 'rm -rf "$TMP"' EXIT
{
  echo '#include "stereo_shim.h"'
  echo 'namespace ORB_SLAM2 {'
  sed -n '37,43p' "$M"
  sed -n '1647,1663p' "$M"
  sed -n '466,640p' "$F"
  sed -n '230,245p' "$F"
  sed -n '382,392p' "$F"
  sed -n '404,434p' "$F"
  sed -n '436,464p' "$F"
  sed -n '327,380p' "$F"
  sed -n '1328,1470p' "$M"
  sed -n '1601,1642p' "$M"
  sed -n '45,137p' "$M"
  sed -n '159,288p' "$M"
  sed -n '405,520p' "$M"
  sed -n '1472,1599p' "$M"
  sed -n '373,383p' "$P"
  sed -n '402,417p' "$P"
  sed -n '269,325p' "$F"
  echo '}'
} > "$TMP/stereo_ref_gen.cpp"
mkdir -p "$OUT"
${CXX:-g++} -std=c++11 -O2 -march=x86-64-v2 -ffp-contract=off -fPIC -shared -w -I"$HERE/shim_stereo" \
    "$TMP/stereo_ref_gen.cpp" "$HERE/stereo_ref_driver.cpp" -o "$OUT/libstereoref.so"
echo "$OUT/libstereoref.so"
