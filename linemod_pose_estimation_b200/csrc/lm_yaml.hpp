// lm_yaml.hpp -- the subset of cv::FileStorage's YAML 1.0 dialect that templates.yml / renderer_params.yml use.
//
// The reference persists its detector with cv::FileStorage (writeLinemod, /root/reference/src/renderer.cpp:56-70;
// readLinemod, src/rgbdDetector.cpp:1668-1680) and its pose table with the same writer (src/renderer.cpp:72-123).
// OpenCV is not available to this library, so this is an independent reader / writer for that dialect:
// "%YAML:1.0" header, 3-space block indentation, "-" block sequences, "[ a, b ]" flow sequences that may wrap,
// "!!opencv-matrix" tags, quoted strings, ".yml.gz" through zlib.  The writer reproduces OpenCV 2.4's layout
// (no "---" document marker); the reader accepts both the 2.4 and the 4.x form.
#pragma once
#include <cstdint>
#include <string>
#include <utility>
#include <vector>

namespace lmyaml {

struct Node {
  enum Kind { NONE, SCALAR, SEQ, MAP };
  Kind kind = NONE;
  std::string sval;                                  // SCALAR text (unquoted / unescaped)
  std::vector<Node> items;                           // SEQ (generic)
  std::vector<double> nums;                          // SEQ of numbers, compact form (items empty)
  bool numeric_seq = false;
  std::vector<std::pair<std::string, Node> > members;  // MAP, file order

  bool empty() const { return kind == NONE; }
  size_t size() const { return kind == SEQ ? (numeric_seq ? nums.size() : items.size()) : (kind == MAP ? members.size() : (kind == SCALAR ? 1 : 0)); }
  const Node& operator[](const std::string& key) const;  // MAP lookup, NONE node when absent
  const Node& at(size_t i) const;                        // generic SEQ element
  double num(size_t i) const;                            // SEQ element as number
  bool as_int(int& v) const;
  bool as_double(double& v) const;
};

// Parses a whole file (plain or .gz).  Returns false and fills err on failure.
bool parse_file(const std::string& path, Node& root, std::string& err);
bool parse_text(const std::string& text, Node& root, std::string& err);

// Streaming writer with cv::FileStorage's operator<< vocabulary.
class Writer {
 public:
  Writer();
  void key(const std::string& k);  // next value / collection is a map member named k
  void begin_map();                // "{"   (block mapping)
  void begin_map_tagged(const std::string& tag);  // "{:tag" e.g. "!!opencv-matrix" (block mapping only)
  void end_map();                  // "}"
  void begin_seq(bool flow);       // "[" (block) or "[:" (flow)
  void end_seq();                  // "]"
  void write_int(int v);
  void write_float(float v);
  void write_double(double v);
  void write_string(const std::string& s);
  const std::string& text();  // finished document
  bool save(const std::string& path, std::string& err);  // plain or .gz by extension

 private:
  struct Frame { bool is_seq; bool flow; bool first; int indent; };
  std::vector<Frame> stack_;
  std::string out_, pending_key_, pending_tag_;
  bool has_key_ = false;
  size_t line_start_ = 0;
  void emit_scalar(const std::string& text);
  void start_collection(bool is_seq, bool flow);
  void newline_indent(int indent);
};

std::string format_float(float v);    // icvFloatToString
std::string format_double(double v);  // icvDoubleToString

}  // namespace lmyaml
