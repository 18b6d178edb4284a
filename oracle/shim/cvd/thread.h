// TEST INFRASTRUCTURE — stand-in for cvd/thread.h: the map maker thread is never started here.
#pragma once
namespace CVD {
class Thread {
 public:
  virtual ~Thread() {}
  void start() {}
  void stop() { stopFlag = true; }
  bool shouldStop() const { return stopFlag; }
  void join() {}
  static void sleep(unsigned) {}
  virtual void run() = 0;
 private:
  bool stopFlag = false;
};
}
