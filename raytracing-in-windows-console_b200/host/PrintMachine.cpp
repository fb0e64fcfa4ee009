#include "PrintMachine.h"

#include <chrono>
#include <cstdio>
#include <cstring>

#include "Timer.h"

int PrintMachine::m_renderingFps = 60;          // reference PrintMachine.cpp:8-9
int PrintMachine::m_printingFps = 60;
size_t PrintMachine::currentWidth = 0;
size_t PrintMachine::currentHeight = 0;
size_t PrintMachine::m_maxSize = 0;
bool PrintMachine::m_running = true;
bool PrintMachine::m_terminateThread = false;
std::unique_ptr<char[]> PrintMachine::m_printBuffer = nullptr;
std::unique_ptr<char[]> PrintMachine::m_backBuffer = nullptr;
size_t PrintMachine::m_printSize = 0;
size_t PrintMachine::m_backBufferPrintSize = 0;
std::string PrintMachine::m_debugInfo = "";
std::mutex PrintMachine::m_backBufferMutex;
bool PrintMachine::m_bShouldSwapBuffer = false;
FILE* PrintMachine::m_sink = nullptr;
bool PrintMachine::m_onlyNewFrames = false;
size_t PrintMachine::m_framesPrinted = 0;
std::thread PrintMachine::m_printThread;
int PrintMachine::m_printingFpsCounter = 0;
float PrintMachine::m_printingFpsTimer = 0.0f;

void PrintMachine::Start(const size_t x, const size_t y)     // reference PrintMachine.cpp:108-152 minus the console set-up
{
    currentWidth = x;
    currentHeight = y;
    m_maxSize = m_charsPerPixel * currentWidth * currentHeight;
    m_printBuffer = std::make_unique<char[]>(m_maxSize);
    m_backBuffer = std::make_unique<char[]>(m_maxSize);
    m_printSize = m_maxSize;
    m_backBufferPrintSize = 0;
    m_running = true;
    m_terminateThread = false;
}

void PrintMachine::CleanUp()
{
    JoinPrintThread();
    m_running = false;
}
bool PrintMachine::CheckIfRunning() { return m_running; }     // reference PrintMachine.cpp:103-106
void PrintMachine::SetDebugInfo(const std::string& s) { m_debugInfo = s; }
void PrintMachine::TerminateThread() { m_terminateThread = true; }

bool PrintMachine::PrintOnce()
{
    bool fresh = false;
    {
        std::lock_guard<std::mutex> g(m_backBufferMutex);
        if (m_bShouldSwapBuffer) {                            // reference :276-285
            m_bShouldSwapBuffer = false;
            m_printSize = m_backBufferPrintSize;
            m_printBuffer.swap(m_backBuffer);
            fresh = true;
        }
    }
    if (m_onlyNewFrames && !fresh) return false;
    FILE* out = m_sink ? m_sink : stdout;
    fputs("\x1b[H", out);                                     // ResetConsolePointer (:287, :308-311)
    fwrite(m_printBuffer.get(), 1, m_printSize, out);          // :288
    fputs("\x1b[m", out);                                     // :295
    fprintf(out, "Rendering FPS: %d    \n", m_renderingFps);  // :296
    fprintf(out, "Printing FPS: %d    \n", m_printingFps);    // :297
    fflush(out);
    ++m_framesPrinted;
    return true;
}

bool PrintMachine::Print()                                    // reference :257-306
{
    Time timer;
    while (!m_terminateThread) {
        timer.Update();
        m_printingFpsCounter++;
        m_printingFpsTimer += static_cast<float>(timer.DeltaTime());
        if (m_printingFpsTimer >= 1.0f) {                     // once every second the fps is updated (:264-270)
            m_printingFps = m_printingFpsCounter;
            m_printingFpsTimer = 0.0f;
            m_printingFpsCounter = 0;
        }
        if (!PrintOnce()) std::this_thread::sleep_for(std::chrono::microseconds(50));
    }
    m_running = false;                                        // :303
    return true;
}

void PrintMachine::StartPrintThread(FILE* sink, bool onlyNewFrames)
{
    JoinPrintThread();
    m_sink = sink;
    m_onlyNewFrames = onlyNewFrames;
    m_terminateThread = false;
    m_running = true;
    m_framesPrinted = 0;
    m_printThread = std::thread([] { PrintMachine::Print(); });   // the reference detaches it (:148-149); here it is joinable
}

void PrintMachine::JoinPrintThread()
{
    if (m_printThread.joinable()) {
        m_terminateThread = true;
        m_printThread.join();
    }
}

size_t PrintMachine::FramesPrinted() { return m_framesPrinted; }

void PrintMachine::UpdateRenderingFPS(const int fps) { m_renderingFps = fps; }
bool PrintMachine::ChangeSize(const size_t x, const size_t y) { currentWidth = x; currentHeight = y; return true; }
const std::mutex* PrintMachine::GetBackBufferMutex() { return &m_backBufferMutex; }
const char* PrintMachine::GetBackBuffer() { return m_backBuffer.get(); }

void PrintMachine::SetDataInBackBuffer(const char* data, const size_t size)   // reference :178-192
{
    std::lock_guard<std::mutex> g(m_backBufferMutex);
    memcpy(m_backBuffer.get(), data, size);
    FlagForBufferSwap();
    SetPrintSize(size);
}

size_t PrintMachine::GetWidth() { return currentWidth; }
size_t PrintMachine::GetHeight() { return currentHeight; }
size_t PrintMachine::GetMaxSize() { return m_maxSize; }
size_t PrintMachine::GetPrintSize() { return m_backBufferPrintSize; }
void PrintMachine::ResetBackBuffer() { memset(m_backBuffer.get(), 0, m_maxSize); }
void PrintMachine::FlagForBufferSwap() { m_bShouldSwapBuffer = true; }
void PrintMachine::SetPrintSize(const size_t n) { m_backBufferPrintSize = n; }
