// oracle/ref_dict_wrap.cpp -- TEST INFRASTRUCTURE.  C entry points around the reference's OWN marker-identification code: Thirdparty/aruco/aruco/
// dictionary.cpp (code tables, loadPredefined, getMarkerImage_id), dictionary_based.cpp (DictionaryBased::setParams / detect: Otsu threshold, cell votes,
// border test, the four rotations, table lookup) and markerlabeler.cpp, compiled unmodified from /root/reference against oracle/arucoshim into
// oracle/_ref/libref_dict.so (oracle/Makefile).  Pins the decode stage of oracle/aruco_oracle.cpp (decode_patch) and the dictionary tables / marker
// rendering of the product (tests/test_oracle_dict_vs_ref.py, tests/golden/dict_ref.npz).
#include <map>
#include <string>
#include "dictionary_based.h"
#include "markermap.h"

namespace aruco {
// Dictionary::createMarkerMap (dictionary.cpp) references MarkerMap members that live in markermap.cpp, which is not compiled (needs FileStorage);
// it is never called here.  Minimal definitions keep the shared object free of unresolved symbols.
MarkerMap::MarkerMap() {}
Marker3DInfo::Marker3DInfo() {}
Marker3DInfo::Marker3DInfo(int _id) : id(_id) {}
std::vector<cv::Point3f> Marker::get3DPoints(float) { return std::vector<cv::Point3f>(); }      // marker.cpp:358, only reached from createMarkerMap
}  // namespace aruco

extern "C" {

// DictionaryBased::detect on one canonical patch (size x size, u8): returns 1 and (id, nRotations) when the reference identifies a marker
int ref_dictionary_detect(const uint8_t* patch, int size, const char* dict_name, int32_t* id, int32_t* nrot) {
    static std::map<std::string, aruco::DictionaryBased*> cache;
    aruco::DictionaryBased*& db = cache[dict_name];
    if (!db) { db = new aruco::DictionaryBased(); db->setParams(aruco::Dictionary::loadPredefined(std::string(dict_name)), 0.f); }
    cv::Mat m(size, size, CV_8UC1);
    for (int y = 0; y < size; y++) memcpy(m.ptr<uchar>(y), patch + (size_t)y * size, size);
    int i = -1, r = -1; std::string info;
    const bool ok = db->detect(m, i, r, info);
    *id = i; *nrot = r;
    return ok ? 1 : 0;
}

// the code table of a predefined dictionary: codes [cap] indexed by id (0 where an id has no entry); returns the highest id + 1, fills nbits and tau
int ref_dictionary_codes(const char* dict_name, uint64_t* codes, int cap, int32_t* nbits, int32_t* tau) {
    aruco::Dictionary d = aruco::Dictionary::loadPredefined(std::string(dict_name));
    *nbits = (int32_t)d.nbits(); *tau = (int32_t)d.tau();
    int n = 0;                                                    // a code listed twice keeps its first id (std::map::insert): the later id has no entry
    for (auto& kv : d.getMapCode()) { if ((int)kv.second < cap) codes[kv.second] = kv.first; if ((int)kv.second + 1 > n) n = (int)kv.second + 1; }
    return n;
}

// Dictionary::getMarkerImage_id(id, bit_size, addWaterMark = false): out [(n + 2) * bit_size]^2, returns the side length
int ref_marker_image(const char* dict_name, int id, int bit_size, uint8_t* out, int cap) {
    aruco::Dictionary d = aruco::Dictionary::loadPredefined(std::string(dict_name));
    cv::Mat img = d.getMarkerImage_id(id, bit_size, false);
    if (img.empty() || img.rows * img.cols > cap) return -1;
    for (int y = 0; y < img.rows; y++) memcpy(out + (size_t)y * img.cols, img.ptr<uchar>(y), img.cols);
    return img.rows;
}

}  // extern "C"
