// Minimal stand-in for <opencv2/core/core.hpp> (OpenCV 2.4 API subset) so that the LINEMOD_B200_WITH_OPENCV branch of
// include/linemod_b200.hpp -- cv::Mat / cv::Rect overloads, Detector::read(FileNode) / readClass / write(FileStorage) /
// writeClass, cv::Ptr modalities, match(..., noArray()) -- is compiled and exercised in CI without OpenCV installed.
// Test infrastructure only.  FileStorage keeps its document as an in-memory tree: a WRITE storage publishes the tree under
// its file name when it is released, a READ storage of the same name finds it (no YAML is parsed or emitted here; the
// on-disk format is the C ABI's job: lm_write_yaml / lm_create_from_yaml, tested against cv2.FileStorage elsewhere).
#ifndef LINEMOD_B200_OPENCV_STUB_CORE_HPP_
#define LINEMOD_B200_OPENCV_STUB_CORE_HPP_

#include <cstddef>
#include <cstring>
#include <map>
#include <memory>
#include <stdexcept>
#include <string>
#include <utility>
#include <vector>

typedef unsigned char uchar;
#define CV_8U 0
#define CV_16U 2
#define CV_MAKETYPE(depth, cn) ((depth) + (((cn)-1) << 3))
#define CV_8UC1 CV_MAKETYPE(CV_8U, 1)
#define CV_8UC3 CV_MAKETYPE(CV_8U, 3)
#define CV_16UC1 CV_MAKETYPE(CV_16U, 1)

namespace cv {

typedef std::string String;  // OpenCV 2.4: cv::String is std::string

template <class T>
class Ptr {  // reference-counted pointer with the converting constructor the reference relies on (Ptr<ColorGradient> -> Ptr<Modality>)
 public:
  Ptr() {}
  Ptr(T* raw) : p_(raw) {}  // NOLINT: implicit like cv::Ptr
  template <class U> Ptr(const Ptr<U>& o) : p_(o.shared()) {}  // NOLINT
  T* operator->() const { return p_.get(); }
  T& operator*() const { return *p_; }
  operator T*() const { return p_.get(); }
  bool empty() const { return !p_; }
  const std::shared_ptr<T>& shared() const { return p_; }

 private:
  std::shared_ptr<T> p_;
};

struct Rect {
  int x, y, width, height;
  Rect() : x(0), y(0), width(0), height(0) {}
  Rect(int x_, int y_, int w, int h) : x(x_), y(y_), width(w), height(h) {}
};

class Mat {
 public:
  struct Step {
    size_t s;
    size_t operator[](int) const { return s; }
    operator size_t() const { return s; }
  };
  uchar* data;
  int rows, cols;
  Step step;
  Mat() : data(nullptr), rows(0), cols(0), type_(CV_8UC1) { step.s = 0; }
  Mat(int r, int c, int type) : rows(r), cols(c), type_(type) {
    step.s = (size_t)c * elemSize();
    buf_.reset(new std::vector<uchar>((size_t)r * step.s, 0));
    data = buf_->data();
  }
  Mat(int r, int c, int type, void* ext, size_t st = 0) : data(static_cast<uchar*>(ext)), rows(r), cols(c), type_(type) {
    step.s = st ? st : (size_t)c * elemSize();
  }
  void create(int r, int c, int type) {
    if (data && r == rows && c == cols && type == type_) return;
    *this = Mat(r, c, type);
  }
  int type() const { return type_; }
  bool empty() const { return data == nullptr || rows == 0 || cols == 0; }
  size_t elemSize() const { return (size_t)(((type_ >> 3) & 7) + 1) * ((type_ & 7) == CV_16U ? 2 : 1); }
  Mat clone() const {
    Mat m(rows, cols, type_);
    for (int y = 0; y < rows; ++y) std::memcpy(m.data + (size_t)y * m.step.s, data + (size_t)y * step.s, (size_t)cols * elemSize());
    return m;
  }

 private:
  int type_;
  std::shared_ptr<std::vector<uchar> > buf_;
};

// OutputArrayOfArrays as Detector::match uses it: noArray() -> not needed; a std::vector<Mat> -> create(n,1,..), create(r,c,type,i), getMat(i)
class _OutputArray {
 public:
  _OutputArray() : v_(nullptr) {}
  _OutputArray(std::vector<Mat>& v) : v_(&v) {}  // NOLINT
  bool needed() const { return v_ != nullptr; }
  void create(int rows, int cols, int type, int i = -1) const {
    if (!v_) return;
    if (i < 0) v_->assign((size_t)rows * (size_t)cols, Mat());
    else (*v_)[(size_t)i] = Mat(rows, cols, type);
  }
  Mat getMat(int i) const { return (*v_)[(size_t)i]; }

 private:
  std::vector<Mat>* v_;
};
typedef const _OutputArray& OutputArrayOfArrays;
inline const _OutputArray& noArray() {
  static const _OutputArray none;
  return none;
}

// ------------------------------------------------------------------------------------------------ FileStorage
namespace stub {
struct Node {
  enum Kind { NONE, INT, REAL, STR, SEQ, MAP } kind;
  long long i;
  double d;
  std::string s;
  std::vector<std::pair<std::string, std::shared_ptr<Node> > > kids;
  Node() : kind(NONE), i(0), d(0) {}
};
inline std::map<std::string, std::shared_ptr<Node> >& registry() {
  static std::map<std::string, std::shared_ptr<Node> > r;
  return r;
}
}  // namespace stub

class FileNodeIterator;
class FileNode {
 public:
  FileNode() {}
  explicit FileNode(const std::shared_ptr<stub::Node>& n) : n_(n) {}
  FileNode operator[](const std::string& key) const {
    if (n_ && n_->kind == stub::Node::MAP)
      for (size_t k = 0; k < n_->kids.size(); ++k)
        if (n_->kids[k].first == key) return FileNode(n_->kids[k].second);
    return FileNode();
  }
  FileNode operator[](const char* key) const { return (*this)[std::string(key)]; }
  FileNode operator[](int i) const {
    if (n_ && n_->kind == stub::Node::SEQ && i >= 0 && (size_t)i < n_->kids.size()) return FileNode(n_->kids[(size_t)i].second);
    return FileNode();
  }
  bool empty() const { return !n_ || n_->kind == stub::Node::NONE; }
  bool isSeq() const { return n_ && n_->kind == stub::Node::SEQ; }
  bool isMap() const { return n_ && n_->kind == stub::Node::MAP; }
  size_t size() const { return n_ && (n_->kind == stub::Node::SEQ || n_->kind == stub::Node::MAP) ? n_->kids.size() : (empty() ? 0 : 1); }
  operator int() const { return !n_ ? 0 : n_->kind == stub::Node::REAL ? (int)(n_->d + (n_->d >= 0 ? 0.5 : -0.5)) : (int)n_->i; }
  operator float() const { return !n_ ? 0.f : n_->kind == stub::Node::INT ? (float)n_->i : (float)n_->d; }
  operator double() const { return !n_ ? 0.0 : n_->kind == stub::Node::INT ? (double)n_->i : n_->d; }
  operator std::string() const { return n_ && n_->kind == stub::Node::STR ? n_->s : std::string(); }
  FileNodeIterator begin() const;
  FileNodeIterator end() const;
  const std::shared_ptr<stub::Node>& node() const { return n_; }

 private:
  std::shared_ptr<stub::Node> n_;
};

class FileNodeIterator {
 public:
  FileNodeIterator(const std::shared_ptr<stub::Node>& n, size_t at) : n_(n), at_(at) {}
  FileNode operator*() const { return FileNode(n_->kids[at_].second); }
  FileNodeIterator& operator++() { ++at_; return *this; }
  bool operator!=(const FileNodeIterator& o) const { return at_ != o.at_ || n_ != o.n_; }
  bool operator==(const FileNodeIterator& o) const { return !(*this != o); }

 private:
  std::shared_ptr<stub::Node> n_;
  size_t at_;
};
inline FileNodeIterator FileNode::begin() const { return FileNodeIterator(n_, 0); }
inline FileNodeIterator FileNode::end() const { return FileNodeIterator(n_, n_ ? n_->kids.size() : 0); }

inline void operator>>(const FileNode& n, std::vector<int>& v) {
  v.clear();
  for (FileNodeIterator it = n.begin(), e = n.end(); it != e; ++it) v.push_back((int)*it);
}
inline void operator>>(const FileNode& n, int& v) { v = (int)n; }
inline void operator>>(const FileNode& n, float& v) { v = (float)n; }
inline void operator>>(const FileNode& n, std::string& v) { v = (std::string)n; }

class FileStorage {
 public:
  enum { READ = 0, WRITE = 1 };
  FileStorage(const std::string& filename, int flags) : name_(filename), write_(flags == WRITE), have_key_(false) {
    if (write_) {
      root_.reset(new stub::Node());
      root_->kind = stub::Node::MAP;
      stack_.push_back(root_);
    } else {
      std::map<std::string, std::shared_ptr<stub::Node> >::iterator it = stub::registry().find(filename);
      if (it != stub::registry().end()) root_ = it->second;
    }
  }
  ~FileStorage() { release(); }
  bool isOpened() const { return (bool)root_; }
  void release() {
    if (write_ && root_) stub::registry()[name_] = root_;
    write_ = false;
  }
  FileNode root() const { return FileNode(root_); }
  FileNode operator[](const std::string& key) const { return root()[key]; }
  FileNode operator[](const char* key) const { return root()[key]; }

  // the streaming writer: in a map strings alternate key / value; "[" "[:" "{" "{:" open a collection, "]" "}" close it
  void put_string(const std::string& s) {
    if (s == "[" || s == "[:") return open(stub::Node::SEQ);
    if (s == "{" || s == "{:") return open(stub::Node::MAP);
    if (s == "]" || s == "}") {
      if (stack_.size() < 2) throw std::runtime_error("FileStorage stub: unbalanced close");
      stack_.pop_back();
      return;
    }
    if (top().kind == stub::Node::MAP && !have_key_) { key_ = s; have_key_ = true; return; }
    std::shared_ptr<stub::Node> n(new stub::Node());
    n->kind = stub::Node::STR; n->s = s;
    add(n);
  }
  void put_int(long long v) { std::shared_ptr<stub::Node> n(new stub::Node()); n->kind = stub::Node::INT; n->i = v; add(n); }
  void put_real(double v) { std::shared_ptr<stub::Node> n(new stub::Node()); n->kind = stub::Node::REAL; n->d = v; add(n); }

 private:
  stub::Node& top() { return *stack_.back(); }
  void add(const std::shared_ptr<stub::Node>& n) {
    if (top().kind == stub::Node::MAP) {
      if (!have_key_) throw std::runtime_error("FileStorage stub: value without a key inside a map");
      top().kids.push_back(std::make_pair(key_, n));
      have_key_ = false;
    } else top().kids.push_back(std::make_pair(std::string(), n));
  }
  void open(stub::Node::Kind k) {
    std::shared_ptr<stub::Node> n(new stub::Node());
    n->kind = k;
    add(n);
    stack_.push_back(n);
  }
  std::string name_;
  bool write_;
  std::shared_ptr<stub::Node> root_;
  std::vector<std::shared_ptr<stub::Node> > stack_;
  std::string key_;
  bool have_key_;
};
inline FileStorage& operator<<(FileStorage& fs, const std::string& s) { fs.put_string(s); return fs; }
inline FileStorage& operator<<(FileStorage& fs, const char* s) { fs.put_string(s); return fs; }
inline FileStorage& operator<<(FileStorage& fs, int v) { fs.put_int(v); return fs; }
inline FileStorage& operator<<(FileStorage& fs, float v) { fs.put_real(v); return fs; }
inline FileStorage& operator<<(FileStorage& fs, double v) { fs.put_real(v); return fs; }
inline FileStorage& operator<<(FileStorage& fs, const std::vector<int>& v) {
  fs.put_string("[:");
  for (size_t i = 0; i < v.size(); ++i) fs.put_int(v[i]);
  fs.put_string("]");
  return fs;
}

}  // namespace cv

#endif  // LINEMOD_B200_OPENCV_STUB_CORE_HPP_
