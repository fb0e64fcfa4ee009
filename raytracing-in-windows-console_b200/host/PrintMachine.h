// PrintMachine.h -- headless mirror of the reference's static frame sink (reference
// PrintMachine.h/.cpp).  Same statics and double-buffer protocol; the Win32 console set-up is gone, the
// frame lands in memory (GetBackBuffer/GetPrintSize), and the reference's print thread (PrintMachine.cpp:
// 138-150 spawns it, :257-306 is its loop) is an opt-in POSIX writer into any FILE* -- a terminal, a pipe or
// a file (StartPrintThread): ESC[H instead of SetConsoleCursorPosition, then fwrite + the two FPS lines.
#pragma once
#include <cstddef>
#include <memory>
#include <cstdio>
#include <mutex>
#include <string>
#include <thread>

#define WIDTHLIMIT 1000
#define HEIGHTLIMIT 500

class PrintMachine
{
protected:
    PrintMachine() = delete;

public:
    static void Start(const size_t x, const size_t y);
    static void CleanUp();
    static bool CheckIfRunning();
    static void SetDebugInfo(const std::string& debugString);
    static void TerminateThread();

    // The reference's print loop (PrintMachine.cpp:257-306): until TerminateThread(), swap buffers if flagged, home the
    // cursor, write the frame and the two FPS lines.  Runs on the caller's thread; StartPrintThread runs it on its own.
    static bool Print();
    // One pass of that loop's body (:274-299).  Returns false if there was no new frame and only new frames are printed.
    static bool PrintOnce();
    // Extensions for headless sinks.  StartPrintThread: what the reference's Start() does at :138-150, into `sink`
    // (default stdout).  onlyNewFrames: skip passes without a freshly swapped frame (the reference re-prints the last
    // frame as fast as the console takes it -- pointless into a file or pipe).
    static void StartPrintThread(FILE* sink = nullptr, bool onlyNewFrames = false);
    static void JoinPrintThread();                 // TerminateThread() + join
    static size_t FramesPrinted();

    static void UpdateRenderingFPS(const int fps);
    static bool ChangeSize(const size_t x, const size_t y);
    static const std::mutex* GetBackBufferMutex();
    static const char* GetBackBuffer();
    static void SetDataInBackBuffer(const char* data, const size_t size);
    static size_t GetWidth();
    static size_t GetHeight();
    static size_t GetMaxSize();
    static size_t GetPrintSize();
    static void ResetBackBuffer();
    static void FlagForBufferSwap();
    static void SetPrintSize(const size_t newSize);

private:
    static int m_renderingFps;
    static int m_printingFps;
    static size_t currentWidth;
    static size_t currentHeight;
    static size_t m_maxSize;
    static bool m_running;
    static bool m_terminateThread;
    static std::unique_ptr<char[]> m_printBuffer;
    static std::unique_ptr<char[]> m_backBuffer;
    static size_t m_printSize;
    static size_t m_backBufferPrintSize;
    static std::string m_debugInfo;
    static std::mutex m_backBufferMutex;
    static bool m_bShouldSwapBuffer;
    static const size_t m_charsPerPixel = 20;   // reference PrintMachine.h:81
    static FILE* m_sink;
    static bool m_onlyNewFrames;
    static size_t m_framesPrinted;
    static std::thread m_printThread;
    static int m_printingFpsCounter;
    static float m_printingFpsTimer;
};
