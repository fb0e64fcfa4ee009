// PrintMachine.h -- headless mirror of the reference's static frame sink (reference
// PrintMachine.h/.cpp).  Same statics and double-buffer protocol; the Win32 console set-up and the
// detached print thread are replaced by an in-memory frame (GetBackBuffer/GetPrintSize) and an
// optional POSIX writer (Print(): ESC[H instead of SetConsoleCursorPosition, then fwrite).
#pragma once
#include <cstddef>
#include <memory>
#include <mutex>
#include <string>

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

    // One pass of the reference's print loop body (PrintMachine.cpp:274-299): swap if flagged,
    // home the cursor, write the frame and the two FPS lines to stdout.
    static bool Print();

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
};
