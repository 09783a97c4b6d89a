// lm_yaml.cpp -- reader / writer for the cv::FileStorage YAML 1.0 dialect (see lm_yaml.hpp).
#include "lm_yaml.hpp"

#include <zlib.h>

#include <cctype>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>

namespace lmyaml {

static const Node kNone;

const Node& Node::operator[](const std::string& key) const {
  if (kind == MAP)
    for (size_t i = 0; i < members.size(); ++i)
      if (members[i].first == key) return members[i].second;
  return kNone;
}
const Node& Node::at(size_t i) const {
  if (kind == SEQ && !numeric_seq && i < items.size()) return items[i];
  return kNone;
}
double Node::num(size_t i) const {
  if (kind != SEQ) return 0;
  if (numeric_seq) return i < nums.size() ? nums[i] : 0;
  double v = 0;
  if (i < items.size()) items[i].as_double(v);
  return v;
}
static bool parse_number(const std::string& s, double& v) {
  if (s.empty()) return false;
  if (s == ".Inf" || s == ".inf") { v = INFINITY; return true; }
  if (s == "-.Inf" || s == "-.inf") { v = -INFINITY; return true; }
  if (s == ".Nan" || s == ".nan" || s == ".NaN") { v = NAN; return true; }
  char* end = nullptr;
  v = std::strtod(s.c_str(), &end);
  return end && *end == '\0' && end != s.c_str();
}
bool Node::as_double(double& v) const { return kind == SCALAR && parse_number(sval, v); }
bool Node::as_int(int& v) const {
  double d;
  if (!as_double(d)) return false;
  v = (int)std::lrint(d);
  return true;
}

// ------------------------------------------------------------------------------------------------ reader
namespace {

struct Line {
  int indent;
  const char* p;  // first non-blank
  const char* e;  // end (comments / trailing blanks stripped)
};

struct Parser {
  std::vector<Line> lines;
  size_t cur = 0;
  std::string err;

  bool fail(const std::string& m) {
    if (err.empty()) err = m + " (line " + std::to_string(cur + 1) + ")";
    return false;
  }

  void split(const std::string& text) {
    const char* s = text.data();
    const char* end = s + text.size();
    while (s < end) {
      const char* nl = (const char*)memchr(s, '\n', end - s);
      const char* le = nl ? nl : end;
      const char* p = s;
      while (p < le && *p == ' ') ++p;
      // strip comment (a '#' outside quotes preceded by blank or at line start) and trailing blanks / CR
      const char* q = p;
      bool inq = false;
      char qc = 0;
      const char* stop = le;
      for (; q < le; ++q) {
        if (inq) { if (*q == '\\' && qc == '"') ++q; else if (*q == qc) inq = false; }
        else if (*q == '"' || *q == '\'') { inq = true; qc = *q; }
        else if (*q == '#' && (q == p || q[-1] == ' ' || q[-1] == '\t')) { stop = q; break; }
      }
      while (stop > p && (stop[-1] == ' ' || stop[-1] == '\r' || stop[-1] == '\t')) --stop;
      if (stop > p) {
        bool directive = (*p == '%') || (stop - p == 3 && !strncmp(p, "---", 3)) || (stop - p == 3 && !strncmp(p, "...", 3));
        if (!directive) { Line l = {(int)(p - s), p, stop}; lines.push_back(l); }
      }
      s = nl ? nl + 1 : end;
    }
  }

  static std::string unquote(const char* p, const char* e) {
    std::string out;
    if (e - p >= 2 && (*p == '"' || *p == '\'') && e[-1] == *p) {
      char qc = *p;
      for (const char* q = p + 1; q < e - 1; ++q) {
        if (*q == '\\' && qc == '"' && q + 1 < e - 1) {
          ++q;
          switch (*q) { case 'n': out += '\n'; break; case 't': out += '\t'; break; case 'r': out += '\r'; break; default: out += *q; }
        } else out += *q;
      }
      return out;
    }
    return std::string(p, e);
  }

  // Flow collection starting at *p ('[' or '{'); may continue over following lines.
  bool parse_flow(const char*& p, const char*& e, Node& n) {
    char open = *p, close = open == '[' ? ']' : '}';
    n.kind = open == '[' ? Node::SEQ : Node::MAP;
    ++p;
    if (p < e && *p == ':') ++p;  // "[:" / "{:" flow markers as written by some emitters
    bool all_num = (open == '[');
    std::vector<std::string> scalars;
    for (;;) {
      while (p < e && (*p == ' ' || *p == ',')) ++p;
      if (p >= e) {  // continue on the next line
        if (++cur >= lines.size()) return fail("unterminated flow collection");
        p = lines[cur].p; e = lines[cur].e;
        continue;
      }
      if (*p == close) { ++p; break; }
      if (*p == '[' || *p == '{') {
        Node child;
        if (!parse_flow(p, e, child)) return false;
        if (open == '[') { all_num = false; n.items.push_back(child); }
        else return fail("nested collection without key in flow map");
        continue;
      }
      // scalar token (possibly "key: value" inside a flow map)
      const char* t = p;
      if (*p == '"' || *p == '\'') {
        char qc = *p++;
        while (p < e && *p != qc) { if (*p == '\\' && qc == '"') ++p; ++p; }
        if (p < e) ++p;
      } else {
        while (p < e && *p != ',' && *p != close && !(open == '{' && *p == ':' && (p + 1 >= e || p[1] == ' '))) ++p;
      }
      const char* te = p;
      while (te > t && te[-1] == ' ') --te;
      if (open == '{') {
        while (p < e && *p == ' ') ++p;
        if (p >= e || *p != ':') return fail("expected ':' in flow map");
        ++p;
        while (p < e && *p == ' ') ++p;
        std::string key = unquote(t, te);
        Node val;
        if (p < e && (*p == '[' || *p == '{')) { if (!parse_flow(p, e, val)) return false; }
        else {
          const char* v = p;
          while (p < e && *p != ',' && *p != close) ++p;
          const char* ve = p;
          while (ve > v && ve[-1] == ' ') --ve;
          val.kind = Node::SCALAR; val.sval = unquote(v, ve);
        }
        n.members.push_back(std::make_pair(key, val));
      } else {
        std::string s = unquote(t, te);
        double d;
        if (all_num && (*t == '"' || *t == '\'' || !parse_number(s, d))) all_num = false;
        Node c; c.kind = Node::SCALAR; c.sval = s;
        n.items.push_back(c);
      }
    }
    if (open == '[' && all_num) {
      n.numeric_seq = true;
      n.nums.reserve(n.items.size());
      for (size_t i = 0; i < n.items.size(); ++i) { double d = 0; parse_number(n.items[i].sval, d); n.nums.push_back(d); }
      n.items.clear();
    }
    return true;
  }

  // Value that starts on the current line at [p, e) (after "key:" or "- "), or on following lines when empty.
  bool parse_value(const char* p, const char* e, int parent_indent, bool after_dash, Node& n) {
    while (p < e && *p == ' ') ++p;
    if (p < e && *p == '!') {  // tag, e.g. !!opencv-matrix
      while (p < e && *p != ' ') ++p;
      while (p < e && *p == ' ') ++p;
    }
    if (p >= e) {
      ++cur;
      if (cur >= lines.size()) { n.kind = Node::NONE; return true; }
      int ind = lines[cur].indent;
      bool dash = *lines[cur].p == '-' && (lines[cur].e - lines[cur].p == 1 || lines[cur].p[1] == ' ');
      if (ind > parent_indent || (ind == parent_indent && dash && !after_dash)) return parse_block(ind, n);
      n.kind = Node::NONE;  // empty value
      return true;
    }
    if (*p == '[' || *p == '{') {
      if (!parse_flow(p, e, n)) return false;
      ++cur;
      return true;
    }
    n.kind = Node::SCALAR;
    n.sval = unquote(p, e);
    ++cur;
    return true;
  }

  static const char* find_key_colon(const char* p, const char* e) {
    if (p < e && (*p == '"' || *p == '\'')) {
      char qc = *p;
      const char* q = p + 1;
      while (q < e && *q != qc) ++q;
      if (q + 1 < e && q[1] == ':') return q + 1;
      if (q + 1 == e) return nullptr;
      return nullptr;
    }
    for (const char* q = p; q < e; ++q)
      if (*q == ':' && (q + 1 == e || q[1] == ' ')) return q;
    return nullptr;
  }

  bool parse_block(int indent, Node& n) {
    if (cur >= lines.size()) { n.kind = Node::NONE; return true; }
    const Line& first = lines[cur];
    bool dash = *first.p == '-' && (first.e - first.p == 1 || first.p[1] == ' ');
    if (dash) {
      n.kind = Node::SEQ;
      while (cur < lines.size() && lines[cur].indent == indent && *lines[cur].p == '-' &&
             (lines[cur].e - lines[cur].p == 1 || lines[cur].p[1] == ' ')) {
        const char* p = lines[cur].p + 1;
        const char* e = lines[cur].e;
        while (p < e && *p == ' ') ++p;
        Node item;
        const char* colon = (p < e && *p != '[' && *p != '{' && *p != '"' ) ? find_key_colon(p, e) : nullptr;
        if (colon) {
          // "- key: value" : a mapping whose first member sits on the dash line
          int child_indent = (int)(p - (lines[cur].p - lines[cur].indent));
          lines[cur].indent = child_indent;
          lines[cur].p = p;
          if (!parse_block(child_indent, item)) return false;
        } else if (!parse_value(p, e, indent, true, item)) return false;
        n.items.push_back(item);
      }
      // compact all-numeric scalar sequences
      bool all_num = !n.items.empty();
      for (size_t i = 0; i < n.items.size() && all_num; ++i) { double d; all_num = n.items[i].kind == Node::SCALAR && parse_number(n.items[i].sval, d); }
      return true;
    }
    if (first.e - first.p == 2 && first.p[0] == '[' && first.p[1] == ']') {  // empty block sequence as cv writes it
      n.kind = Node::SEQ; ++cur; return true;
    }
    n.kind = Node::MAP;
    while (cur < lines.size() && lines[cur].indent == indent) {
      const char* p = lines[cur].p;
      const char* e = lines[cur].e;
      if (*p == '-' && (e - p == 1 || p[1] == ' ')) break;
      const char* colon = find_key_colon(p, e);
      if (!colon) return fail("expected 'key:'");
      std::string key = unquote(p, colon);
      Node val;
      if (!parse_value(colon + 1, e, indent, false, val)) return false;
      n.members.push_back(std::make_pair(key, val));
    }
    if (cur < lines.size() && lines[cur].indent > indent) return fail("unexpected indentation");
    return true;
  }
};

bool read_all(const std::string& path, std::string& text, std::string& err) {
  gzFile f = gzopen(path.c_str(), "rb");  // transparently reads plain files too
  if (!f) { err = "cannot open " + path; return false; }
  char buf[1 << 16];
  int n;
  while ((n = gzread(f, buf, sizeof(buf))) > 0) text.append(buf, n);
  gzclose(f);
  if (n < 0) { err = "read error on " + path; return false; }
  return true;
}

}  // namespace

bool parse_text(const std::string& text, Node& root, std::string& err) {
  Parser ps;
  ps.split(text);
  root = Node();
  if (ps.lines.empty()) { root.kind = Node::MAP; return true; }
  if (!ps.parse_block(ps.lines[0].indent, root)) { err = ps.err; return false; }
  if (ps.cur < ps.lines.size()) { err = "trailing content at line " + std::to_string(ps.cur + 1); return false; }
  return true;
}

bool parse_file(const std::string& path, Node& root, std::string& err) {
  std::string text;
  if (!read_all(path, text, err)) return false;
  if (!parse_text(text, root, err)) { err = path + ": " + err; return false; }
  return true;
}

// ------------------------------------------------------------------------------------------------ writer
std::string format_float(float v) {
  char buf[64];
  if (std::isfinite(v)) {
    int iv = (int)std::lrintf(v);
    if ((float)iv == v) snprintf(buf, sizeof(buf), "%d.", iv);
    else snprintf(buf, sizeof(buf), "%.8e", (double)v);
  } else if (std::isnan(v)) snprintf(buf, sizeof(buf), ".Nan");
  else snprintf(buf, sizeof(buf), v < 0 ? "-.Inf" : ".Inf");
  return buf;
}
std::string format_double(double v) {
  char buf[64];
  if (std::isfinite(v)) {
    int iv = (int)std::lrint(v);
    if ((double)iv == v) snprintf(buf, sizeof(buf), "%d.", iv);
    else snprintf(buf, sizeof(buf), "%.16e", v);
  } else if (std::isnan(v)) snprintf(buf, sizeof(buf), ".Nan");
  else snprintf(buf, sizeof(buf), v < 0 ? "-.Inf" : ".Inf");
  return buf;
}

Writer::Writer() {
  out_ = "%YAML:1.0\n";
  line_start_ = out_.size();
  Frame root = {false, false, true, 0};
  stack_.push_back(root);
}
void Writer::key(const std::string& k) { pending_key_ = k; has_key_ = true; }

void Writer::newline_indent(int indent) {
  if (out_.size() > line_start_) out_ += '\n';
  line_start_ = out_.size();
  out_.append((size_t)indent, ' ');
}

static const int kWrapMargin = 71;

void Writer::emit_scalar(const std::string& text) {
  Frame& f = stack_.back();
  if (f.flow) {
    if (!f.first) out_ += ',';
    int keylen = has_key_ ? (int)pending_key_.size() + 1 : 0;
    int new_offset = (int)(out_.size() - line_start_) + keylen + (int)text.size();
    if (new_offset > kWrapMargin && new_offset - f.indent > 10) newline_indent(f.indent);
    else out_ += ' ';
    if (has_key_) { out_ += pending_key_; out_ += ':'; }
    out_ += text;
  } else {
    newline_indent(f.indent);
    if (f.is_seq) { out_ += "- "; out_ += text; }
    else { out_ += pending_key_; out_ += ": "; out_ += text; }
  }
  f.first = false;
  has_key_ = false;
}

void Writer::start_collection(bool is_seq, bool flow) {
  Frame& p = stack_.back();
  Frame f = {is_seq, flow || p.flow, true, 0};
  if (p.flow) {
    if (!p.first) out_ += ',';
    out_ += ' ';
    if (has_key_) { out_ += pending_key_; out_ += ':'; }
    out_ += is_seq ? "[" : "{";
    f.indent = p.indent;
  } else {
    newline_indent(p.indent);
    if (p.is_seq) out_ += "-";
    else { out_ += pending_key_; out_ += ":"; }
    if (!pending_tag_.empty()) { out_ += " "; out_ += pending_tag_; pending_tag_.clear(); }
    if (f.flow) { out_ += is_seq ? " [" : " {"; f.indent = p.indent + 4; }
    else f.indent = p.indent + 3;
  }
  p.first = false;
  has_key_ = false;
  stack_.push_back(f);
}
void Writer::begin_map() { start_collection(false, false); }
void Writer::begin_map_tagged(const std::string& tag) { pending_tag_ = tag; start_collection(false, false); }
void Writer::begin_seq(bool flow) { start_collection(true, flow); }
void Writer::end_map() {
  Frame f = stack_.back();
  stack_.pop_back();
  if (f.flow) out_ += " }";
}
void Writer::end_seq() {
  Frame f = stack_.back();
  stack_.pop_back();
  if (f.flow) out_ += " ]";
  else if (f.first) { newline_indent(f.indent); out_ += "[]"; }  // empty block sequence, as cv writes it
}
void Writer::write_int(int v) { emit_scalar(std::to_string(v)); }
void Writer::write_float(float v) { emit_scalar(format_float(v)); }
void Writer::write_double(double v) { emit_scalar(format_double(v)); }
void Writer::write_string(const std::string& s) {
  // icvYMLWriteString quoting rules
  size_t len = s.size();
  bool need_quote = len == 0 || s[0] == ' ';
  std::string body;
  for (size_t i = 0; i < len; ++i) {
    char c = s[i];
    if (!need_quote && !isalnum((unsigned char)c) && c != '_' && c != ' ' && c != '-' && c != '(' && c != ')' &&
        c != '/' && c != '+' && c != ';')
      need_quote = true;
    if (!isalnum((unsigned char)c) && (!isprint((unsigned char)c) || c == '\\' || c == '\'' || c == '"')) {
      body += '\\';
      if (isprint((unsigned char)c)) body += c;
      else if (c == '\n') body += 'n';
      else if (c == '\r') body += 'r';
      else if (c == '\t') body += 't';
      else { char b[8]; snprintf(b, sizeof(b), "x%02x", (unsigned char)c); body += b; }
    } else body += c;
  }
  if (!need_quote && len && (isdigit((unsigned char)s[0]) || s[0] == '+' || s[0] == '-' || s[0] == '.')) need_quote = true;
  emit_scalar(need_quote ? "\"" + body + "\"" : body);
}
const std::string& Writer::text() {
  if (out_.empty() || out_[out_.size() - 1] != '\n') out_ += '\n';
  line_start_ = out_.size();
  return out_;
}
bool Writer::save(const std::string& path, std::string& err) {
  const std::string& t = text();
  bool gz = path.size() > 3 && path.compare(path.size() - 3, 3, ".gz") == 0;
  if (gz) {
    gzFile f = gzopen(path.c_str(), "wb");
    if (!f) { err = "cannot open " + path; return false; }
    bool ok = gzwrite(f, t.data(), (unsigned)t.size()) == (int)t.size();
    gzclose(f);
    if (!ok) { err = "write error on " + path; return false; }
    return true;
  }
  FILE* f = fopen(path.c_str(), "wb");
  if (!f) { err = "cannot open " + path; return false; }
  bool ok = fwrite(t.data(), 1, t.size(), f) == t.size();
  fclose(f);
  if (!ok) { err = "write error on " + path; return false; }
  return true;
}

}  // namespace lmyaml
