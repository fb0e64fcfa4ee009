#include "PrintMachine.h"

#include <cstdio>
#include <cstring>

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

void PrintMachine::CleanUp() { m_running = false; }
bool PrintMachine::CheckIfRunning() { return m_running && !m_terminateThread; }
void PrintMachine::SetDebugInfo(const std::string& s) { m_debugInfo = s; }
void PrintMachine::TerminateThread() { m_terminateThread = true; }

bool PrintMachine::Print()
{
    {
        std::lock_guard<std::mutex> g(m_backBufferMutex);
        if (m_bShouldSwapBuffer) {                            // reference :276-285
            m_bShouldSwapBuffer = false;
            m_printSize = m_backBufferPrintSize;
            m_printBuffer.swap(m_backBuffer);
        }
    }
    fputs("\x1b[H", stdout);                                  // ResetConsolePointer
    fwrite(m_printBuffer.get(), 1, m_printSize, stdout);
    printf("\x1b[m");
    printf("Rendering FPS: %d    \n", m_renderingFps);         // reference :297-299
    printf("Printing FPS: %d    \n", m_printingFps);
    return true;
}

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
