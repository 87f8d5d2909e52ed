// DumpGoldens — run ON A MAC inside the FluidAudio package to pin what this repository cannot pin by itself:
// the VALUES Apple's closed code (vDSP_DFT_zop, vDSP_mmul, vForce, cblas_dgemm) produces for
//   AudioMelSpectrogram.computeFlatTransposed / computeFlat / compute   (Shared/AudioMelSpectrogram.swift:132-456)
//   VBxClustering.refine                                                 (Diarizer/Offline/Clustering/VBxClustering.swift:41-165)
//   AudioConverter.resample (AVAudioConverter, Mastering / max quality)   (Shared/AudioConverter.swift:60-71, 299-375)
// on the synthetic fixtures `python tests/golden/swift_fixtures.py inputs <dir>` writes.
//
// NOT compiled in this repository (no Swift toolchain in the build image).  To use it, in a FluidAudio checkout:
//   1. copy this file to Sources/FluidAudioCLI/Commands/DumpGoldens.swift (or any executable target that can
//      `@testable import FluidAudio`: VBxClustering is internal) and call `DumpGoldens.run(inputs:, outputs:)`
//      from the CLI's argument switch;
//   2. swift run -c release fluidaudiocli dump-goldens <inputs dir> <outputs dir>
//   3. bring <outputs dir> back and run `python tests/golden/swift_fixtures.py pack <outputs dir>`:
//      it writes tests/golden/swift_mel.npz / swift_vbx.npz / swift_resample.npz, which tests/test_swift_goldens.py
//      (oracle, CPU) and tests/test_gpu_parity.py::test_swift_goldens_when_present (CUDA path) consume.
// File formats are raw little-endian arrays; shapes travel in manifest.json next to them.
import CoreML
import Foundation

@testable import FluidAudio

@available(macOS 14.0, iOS 17.0, *)
enum DumpGoldens {
    static func readFloats(_ url: URL) throws -> [Float] {
        let d = try Data(contentsOf: url)
        return d.withUnsafeBytes { Array($0.bindMemory(to: Float.self)) }
    }
    static func readDoubles(_ url: URL) throws -> [Double] {
        let d = try Data(contentsOf: url)
        return d.withUnsafeBytes { Array($0.bindMemory(to: Double.self)) }
    }
    static func readInts(_ url: URL) throws -> [Int32] {
        let d = try Data(contentsOf: url)
        return d.withUnsafeBytes { Array($0.bindMemory(to: Int32.self)) }
    }
    static func write<T>(_ values: [T], _ url: URL) throws {
        try values.withUnsafeBufferPointer { Data(buffer: $0) }.write(to: url)
    }

    static func run(inputs: URL, outputs: URL) async throws {
        try FileManager.default.createDirectory(at: outputs, withIntermediateDirectories: true)
        var manifest: [String: Any] = ["host": ProcessInfo.processInfo.operatingSystemVersionString]

        // ---- mel: every entry point on the tone+noise and the speech-like fixture, 80 and 128 mels ----
        for name in ["tone_noise", "speech_like"] {
            let audio = try readFloats(inputs.appendingPathComponent("audio_\(name).f32"))
            for nMels in [80, 128] {
                let mel = AudioMelSpectrogram(nMels: nMels)
                let t = mel.computeFlatTransposed(audio: audio, lastAudioSample: 0, paddingMode: .center)
                try write(t.mel, outputs.appendingPathComponent("mel_\(name)_\(nMels)_center.f32"))
                let p = mel.computeFlatTransposed(
                    audio: audio, lastAudioSample: 0.25, paddingMode: .prePadded, expectedFrameCount: nil)
                try write(p.mel, outputs.appendingPathComponent("mel_\(name)_\(nMels)_prepadded.f32"))
                let f = mel.computeFlat(audio: audio, lastAudioSample: 0)
                try write(f.mel, outputs.appendingPathComponent("mel_\(name)_\(nMels)_flat.f32"))
                manifest["mel_\(name)_\(nMels)"] = [
                    "center": [t.melLength, t.numFrames], "prepadded": [p.melLength, p.numFrames],
                    "flat": [f.melLength, f.numFrames],
                ]
            }
            let legacy = AudioMelSpectrogram(nMels: 128).compute(audio: audio)
            try write(legacy.mel[0].flatMap { $0 }, outputs.appendingPathComponent("mel_\(name)_128_legacy.f32"))
            manifest["mel_\(name)_128_legacy"] = [legacy.melLength]
        }
        try write(AudioMelSpectrogram(nMels: 80).getFilterbank(), outputs.appendingPathComponent("filterbank_80.f32"))
        try write(AudioMelSpectrogram(nMels: 80).getHannWindow(), outputs.appendingPathComponent("hann_400.f32"))

        // ---- AudioConverter: 48 kHz / 44.1 kHz / 8 kHz mono -> 16 kHz through AVAudioConverter ----
        let conv = AudioConverter()
        for rate in [48000, 44100, 8000] {
            let x = try readFloats(inputs.appendingPathComponent("pcm_\(rate).f32"))
            let y = try conv.resample(x, from: Double(rate))
            try write(y, outputs.appendingPathComponent("resampled_\(rate).f32"))
            manifest["resampled_\(rate)"] = [y.count]
        }

        // ---- VBx: refine() on the synthetic rho / psi / AHC labels ----
        let dims = try readInts(inputs.appendingPathComponent("vbx_shape.i32"))   // [T, D]
        let T = Int(dims[0])
        let D = Int(dims[1])
        let rhoFlat = try readDoubles(inputs.appendingPathComponent("vbx_rho.f64"))
        let psi = try readDoubles(inputs.appendingPathComponent("vbx_psi.f64"))
        let initial = try readInts(inputs.appendingPathComponent("vbx_initial.i32")).map { Int($0) }
        let rho = (0..<T).map { Array(rhoFlat[$0 * D..<($0 + 1) * D]) }
        // PLDATransform needs an MLModel only for transform(); refine() reads phiParameters alone.
        let models = try await OfflineDiarizerModels.load()
        let plda = PLDATransform(pldaRhoModel: models.pldaRhoModel, psi: psi)
        let out = VBxClustering(config: .default, pldaTransform: plda).refine(rhoFeatures: rho, initialClusters: initial)
        try write(out.gamma.flatMap { $0 }, outputs.appendingPathComponent("vbx_gamma.f64"))
        try write(out.pi, outputs.appendingPathComponent("vbx_pi.f64"))
        try write(out.elbos, outputs.appendingPathComponent("vbx_elbos.f64"))
        try write(out.hardClusters.flatMap { $0 }.map { Int32($0) }, outputs.appendingPathComponent("vbx_hard.i32"))
        manifest["vbx"] = ["T": T, "D": D, "S": out.numClusters, "iterations": out.elbos.count]

        let json = try JSONSerialization.data(withJSONObject: manifest, options: [.prettyPrinted, .sortedKeys])
        try json.write(to: outputs.appendingPathComponent("manifest.json"))
    }
}
