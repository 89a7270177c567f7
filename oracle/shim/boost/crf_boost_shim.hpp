// =============================================================================
// oracle/shim/boost/crf_boost_shim.hpp — TEST INFRASTRUCTURE ONLY.
//
// A minimal stand-in for the Boost 1.55 pieces the reference's inference path uses, so that the UNMODIFIED
// reference sources compile in an image without Boost headers (oracle/Makefile target _ref/libcrf_ref.so).
// Every <boost/...> header under oracle/shim/boost/ forwards here.
//
//   * boost::archive::text_iarchive — a reader of the Boost text-archive v10 grammar that the reference's own
//     serialize() methods drive (Tree.hpp:334-343, TreeNode.hpp:148-164, ThresholdSplit.hpp:60-67,
//     ImageSample.hpp:83-90, Constants.hpp:44-59, opencv_serialization.hpp:65-79, HeadPoseSample.hpp:154-161,
//     MPSample.hpp:149-158).  Rules (checked on all 115 shipped files): whitespace-separated
//     tokens; a class type emits `tracking version` the first time it appears in a file; a pointer emits
//     `class_id [tracking version] object_id` (class_id -1 = NULL); std::string = `length chars`;
//     std::vector<T> = `count item_version items` (a vector of class type is itself a class type).
//   * boost::asio::io_service / thread_group / bind / thread — what include/ThreadPool.hpp needs, on std::thread.
//   * filesystem, iostreams, lexical_cast, split/is_any_of, numeric::bounds, random — thin std:: equivalents.
//   * text_oarchive — compile-only (Tree::save is training code).
// =============================================================================
#ifndef CRF_BOOST_SHIM_HPP
#define CRF_BOOST_SHIM_HPP

#include <algorithm>
#include <condition_variable>
#include <cstdlib>
#include <ctime>
#include <deque>
#include <dirent.h>
#include <fstream>
#include <functional>
#include <istream>
#include <limits>
#include <memory>
#include <mutex>
#include <ostream>
#include <random>
#include <set>
#include <sstream>
#include <stdexcept>
#include <string>
#include <sys/stat.h>
#include <thread>
#include <type_traits>
#include <typeindex>
#include <vector>

namespace boost {

// ---- smart pointers / bind -------------------------------------------------------------------
template <typename T> using shared_ptr = std::shared_ptr<T>;
template <typename... A> inline auto bind(A&&... a) -> decltype(std::bind(std::forward<A>(a)...)) { return std::bind(std::forward<A>(a)...); }

// ---- thread ----------------------------------------------------------------------------------
class thread {
 public:
  static unsigned hardware_concurrency();   // oracle/shim/cvshim.cc: std::thread::hardware_concurrency(), or CRF_REF_THREADS
};
class thread_group {
 public:
  ~thread_group() { join_all(); }
  template <typename F> void create_thread(F f) { threads_.emplace_back(f); }
  void join_all() { for (auto& t : threads_) if (t.joinable()) t.join(); threads_.clear(); }
 private:
  std::vector<std::thread> threads_;
};

// ---- asio::io_service: post / run / work / stop as include/ThreadPool.hpp:24-64 uses them ------
namespace asio {
class io_service {
 public:
  class work {
   public:
    explicit work(io_service& s) : s_(s) { std::lock_guard<std::mutex> lk(s_.m_); s_.work_++; }
    ~work() { { std::lock_guard<std::mutex> lk(s_.m_); s_.work_--; } s_.cv_.notify_all(); }
   private:
    io_service& s_;
  };
  io_service() {}
  explicit io_service(size_t) {}
  template <typename F> void post(F f) { { std::lock_guard<std::mutex> lk(m_); q_.emplace_back(std::move(f)); } cv_.notify_one(); }
  // run(): executes handlers until there is no more work (no queued handler and no io_service::work alive) or stop()
  size_t run() {
    size_t n = 0;
    for (;;) {
      std::function<void()> f;
      {
        std::unique_lock<std::mutex> lk(m_);
        cv_.wait(lk, [this] { return stopped_ || !q_.empty() || work_ == 0; });
        if (stopped_) return n;
        if (q_.empty()) return n;   // work_ == 0 and nothing queued
        f = std::move(q_.front());
        q_.pop_front();
      }
      f();
      n++;
    }
  }
  void stop() { { std::lock_guard<std::mutex> lk(m_); stopped_ = true; } cv_.notify_all(); }
 private:
  friend class work;
  std::mutex m_;
  std::condition_variable cv_;
  std::deque<std::function<void()>> q_;
  int work_ = 0;
  bool stopped_ = false;
};
}  // namespace asio

// ---- numeric::bounds ---------------------------------------------------------------------------
namespace numeric {
template <typename T> struct bounds {
  static T lowest() { return std::numeric_limits<T>::lowest(); }
  static T highest() { return std::numeric_limits<T>::max(); }
};
}  // namespace numeric

// ---- random (training code only) ---------------------------------------------------------------
typedef std::mt19937 mt19937;
template <typename T = int> class uniform_int {
 public:
  uniform_int(T lo, T hi) : lo_(lo), hi_(hi) {}
  T lo_, hi_;
};
template <typename Engine, typename Dist> class variate_generator {
 public:
  variate_generator(Engine e, Dist d) : e_(e), d_(d) {}
  int operator()() { std::uniform_int_distribution<int> u((int)d_.lo_, (int)std::max(d_.lo_, d_.hi_)); return u(e_); }
 private:
  Engine e_;
  Dist d_;
};

// ---- lexical_cast / string algorithms ----------------------------------------------------------
struct bad_lexical_cast : std::runtime_error { bad_lexical_cast() : std::runtime_error("bad lexical cast") {} };
template <typename T> inline T lexical_cast(const std::string& s) {
  std::istringstream is(s);
  T v;
  if (!(is >> v)) throw bad_lexical_cast();
  return v;
}
struct is_any_of { explicit is_any_of(const std::string& s) : set(s) {} std::string set; };
inline void split(std::vector<std::string>& out, const std::string& in, const is_any_of& sep) {
  out.clear();
  std::string cur;
  for (char ch : in) {
    if (sep.set.find(ch) != std::string::npos) { out.push_back(cur); cur.clear(); }
    else cur.push_back(ch);
  }
  out.push_back(cur);
}

// ---- filesystem --------------------------------------------------------------------------------
namespace filesystem {
struct file_status { bool dir = false; };
class path {
 public:
  path() {}
  path(const std::string& s) : s_(s) {}
  path(const char* s) : s_(s) {}
  const std::string& string() const { return s_; }
 private:
  std::string s_;
};
inline bool exists(const path& p) { struct stat st; return ::stat(p.string().c_str(), &st) == 0; }
inline bool is_directory(const file_status& s) { return s.dir; }
class directory_entry {
 public:
  const filesystem::path& path() const { return p_; }
  file_status status() const { struct stat st; file_status s; s.dir = ::stat(p_.string().c_str(), &st) == 0 && S_ISDIR(st.st_mode); return s; }
  filesystem::path p_;
};
class directory_iterator {
 public:
  directory_iterator() {}
  explicit directory_iterator(const filesystem::path& dir) {
    list_ = std::make_shared<std::vector<directory_entry>>();
    if (DIR* d = ::opendir(dir.string().c_str())) {
      while (dirent* e = ::readdir(d)) {
        const std::string n = e->d_name;
        if (n == "." || n == "..") continue;
        directory_entry de; de.p_ = filesystem::path(dir.string() + "/" + n);
        list_->push_back(de);
      }
      ::closedir(d);
    }
    if (list_->empty()) list_.reset();
  }
  directory_iterator& operator++() { if (list_ && ++i_ >= list_->size()) list_.reset(); return *this; }
  const directory_entry& operator*() const { return (*list_)[i_]; }
  const directory_entry* operator->() const { return &(*list_)[i_]; }
  bool operator!=(const directory_iterator& o) const { return !(list_ == o.list_ && (!list_ || i_ == o.i_)); }
 private:
  std::shared_ptr<std::vector<directory_entry>> list_;
  size_t i_ = 0;
};
}  // namespace filesystem

// ---- iostreams::stream<file_source> -------------------------------------------------------------
namespace iostreams {
struct file_source {};
template <typename Device> class stream : public std::ifstream {
 public:
  explicit stream(const char* p) : std::ifstream(p) {}
  explicit stream(const std::string& p) : std::ifstream(p.c_str()) {}
};
}  // namespace iostreams

// ---- serialization -----------------------------------------------------------------------------
namespace serialization {
class access {
 public:
  template <class Archive, class T> static void serialize(Archive& ar, T& t, unsigned version) { t.serialize(ar, version); }
};
// third argument of serialize(): a type of THIS namespace, so that the reference's free overloads for cv::Rect_ / cv::Point_
// (include/opencv_serialization.hpp:65-79, declared in boost::serialization after this header) are found by ADL
struct version_type { unsigned v; operator unsigned() const { return v; } };
struct binary_object { void* p; size_t n; };
inline binary_object make_binary_object(void* p, size_t n) { binary_object b; b.p = p; b.n = n; return b; }
// free serialize(): classes with a member serialize() go through access; cv::Rect_ / cv::Point_ have free overloads
// (include/opencv_serialization.hpp:65-79) which are found by ADL-free ordinary lookup at instantiation below.
template <class Archive, class T> inline void serialize(Archive& ar, T& t, const unsigned version) { access::serialize(ar, t, version); }
}  // namespace serialization
#define BOOST_SERIALIZATION_SPLIT_FREE(T) static_assert(true, "split_free: cv::Mat archives are not on the inference path")

namespace archive {
class archive_exception : public std::exception {
 public:
  explicit archive_exception(const std::string& s) : s_(s) {}
  const char* what() const noexcept override { return s_.c_str(); }
 private:
  std::string s_;
};

class text_iarchive {
 public:
  explicit text_iarchive(std::istream& is) : is_(is) {
    std::string sig;
    load(sig);
    if (sig != "serialization::archive") throw archive_exception("invalid signature");
    load(lib_version_);
  }
  template <class T> text_iarchive& operator>>(T& t) { return *this & t; }
  template <class T> text_iarchive& operator&(T& t) { load(t); return *this; }
  text_iarchive& operator&(const serialization::binary_object&) { throw archive_exception("binary objects are not supported by the shim"); }

 private:
  template <class T> typename std::enable_if<std::is_arithmetic<T>::value>::type load(T& v) {
    if (!(is_ >> v)) throw archive_exception("input stream error");
  }
  void load(bool& v) { int i = 0; if (!(is_ >> i)) throw archive_exception("input stream error"); v = i != 0; }
  void load(std::string& s) {
    size_t n = 0;
    if (!(is_ >> n)) throw archive_exception("input stream error");
    is_.get();   // the single separating space
    s.resize(n);
    if (n) is_.read(&s[0], (std::streamsize)n);
    if (!is_) throw archive_exception("input stream error");
  }
  // first occurrence of a class type in this archive: `tracking_level version`
  template <class T> unsigned class_info() {
    const std::type_index id(typeid(T));
    auto it = std::find_if(seen_.begin(), seen_.end(), [&](const std::pair<std::type_index, unsigned>& p) { return p.first == id; });
    if (it != seen_.end()) return it->second;
    int tracking = 0; unsigned version = 0;
    load(tracking); load(version);
    seen_.emplace_back(id, version);
    return version;
  }
  template <class T> typename std::enable_if<std::is_class<T>::value>::type load(T& t) {
    const unsigned version = class_info<T>();
    serialize(*this, t, boost::serialization::version_type{version});
  }
  template <class T> void load(std::vector<T>& v) {
    if (std::is_class<T>::value) class_info<std::vector<T>>();
    size_t count = 0; unsigned item_version = 0;
    load(count);
    if (lib_version_ > 3) load(item_version);
    v.clear();
    v.resize(count);
    for (size_t i = 0; i < count; i++) load(v[i]);
  }
  template <class T> void load(T*& p) {
    int class_id = 0;
    load(class_id);
    if (class_id == -1) { p = nullptr; return; }
    class_info<T>();
    unsigned object_id = 0;
    load(object_id);
    p = new T();
    serialize(*this, *p, boost::serialization::version_type{0u});
  }
  std::istream& is_;
  unsigned lib_version_ = 0;
  std::vector<std::pair<std::type_index, unsigned>> seen_;
};

// compile-only: Tree::save / oa << *this are training code
class text_oarchive {
 public:
  explicit text_oarchive(std::ostream&) {}
  template <class T> text_oarchive& operator<<(const T&) { throw archive_exception("text_oarchive is a compile-only stub"); }
  template <class T> text_oarchive& operator&(const T&) { throw archive_exception("text_oarchive is a compile-only stub"); }
};
}  // namespace archive
}  // namespace boost
#endif
