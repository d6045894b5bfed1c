// stand-in (oracle/cvstub, TEST INFRASTRUCTURE): declarations of cv::FileStorage / cv::FileNode so that the YAML
// save / load members of the reference's DBoW2 vocabulary template parse.  They are virtual, hence instantiated, but
// never CALLED by the oracle (the vocabulary is loaded with the reference's loadFromTextFile); every operation is inert.
#pragma once
#include <sstream>
#include <string>
namespace cv {
class FileNode {
public:
    FileNode operator[](const std::string&) const { return FileNode(); }
    FileNode operator[](const char*) const { return FileNode(); }
    FileNode operator[](int) const { return FileNode(); }
    size_t size() const { return 0; }
    operator int() const { return 0; }
    operator double() const { return 0.0; }
    operator float() const { return 0.f; }
    operator std::string() const { return std::string(); }
};
class FileStorage {
public:
    enum { READ = 0, WRITE = 1 };
    FileStorage() {}
    FileStorage(const char*, int) {}
    FileStorage(const std::string&, int) {}
    bool isOpened() const { return false; }
    void release() {}
    FileNode operator[](const std::string&) const { return FileNode(); }
    FileNode operator[](const char*) const { return FileNode(); }
};
template <class T> inline FileStorage& operator<<(FileStorage& fs, const T&) { return fs; }
}  // namespace cv
