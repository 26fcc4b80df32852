// oracle/vocshim/opencv2/core/core.hpp -- TEST INFRASTRUCTURE.  What the reference's vendored DBoW2 needs from OpenCV to compile unmodified into
// oracle/_ref/libref_voc.so: the cv::Mat stand-in of oracle/matchshim plus a cv::FileStorage / cv::FileNode that only throws:
// TemplatedVocabulary.h names them in its YAML save / load members, which are never called here (vocabularies are loaded with loadFromTextFile).
#pragma once
#include <sstream>
#include <stdexcept>
#include <string>
#include "../../../matchshim/opencv2/core/core.hpp"

namespace cv {

// save / load are virtual, so their bodies are instantiated with the vtable: every member exists and refuses to work
struct NoFileStorage : std::runtime_error { NoFileStorage() : std::runtime_error("cv::FileStorage is not part of the stand-in (use loadFromTextFile)") {} };
class FileNode {
public:
    FileNode operator[](const char*) const { throw NoFileStorage(); }
    FileNode operator[](const std::string&) const { throw NoFileStorage(); }
    FileNode operator[](int) const { throw NoFileStorage(); }
    size_t size() const { throw NoFileStorage(); }
    operator int() const { throw NoFileStorage(); }
    operator double() const { throw NoFileStorage(); }
    operator std::string() const { throw NoFileStorage(); }
};
class FileStorage {
public:
    enum { READ = 0, WRITE = 1 };
    FileStorage(const char*, int) { throw NoFileStorage(); }
    FileStorage(const std::string&, int) { throw NoFileStorage(); }
    bool isOpened() const { return false; }
    void release() {}
    FileNode operator[](const char*) const { throw NoFileStorage(); }
    FileNode operator[](const std::string&) const { throw NoFileStorage(); }
};
template <typename T> FileStorage& operator<<(FileStorage&, const T&) { throw NoFileStorage(); }

}  // namespace cv
