#pragma once
#include <chrono>
namespace CVD {
struct cvd_timer { double get_time() const { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); } };
static cvd_timer timer;
}
