// Timer.h -- reference Timer.h/Timer.cpp restated (high_resolution_clock, long double seconds).
#pragma once
#include <chrono>

class Time
{
    using t_clock = std::chrono::high_resolution_clock;
    using t_dSec = std::chrono::duration<long double, std::ratio<1, 1>>;

public:
    Time() : m_start(t_clock::now()), m_loopBegin(t_clock::now()), m_deltaTime(0) {}
    virtual ~Time() = default;
    long double SinceStart() { return t_dSec(t_clock::now() - m_start).count(); }
    long double DeltaTime() { return m_deltaTime.count(); }
    void Update()
    {
        const auto now = t_clock::now();
        m_deltaTime = now - m_loopBegin;
        m_loopBegin = now;
    }

private:
    t_clock::time_point m_start;
    t_clock::time_point m_loopBegin;
    t_dSec m_deltaTime;
};
