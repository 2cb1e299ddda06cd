// cli_main.cpp — a C++ host above the C-ABI, the way Sailor's own executable would call the path tracer.
//
// The reference builds PathTracer::Params from the command line (reference Raytracing/PathTracer.cpp:30-73) and calls
// PathTracer::Run (:75-575); nothing in the engine does so today (Runtime/Sailor.cpp only includes the header).
// This is that missing caller, against include/sailor_pt.h only (no CUDA, no torch in sight): the same flags, the same
// defaults (PathTracer.h:21-32), plus the extensions of the drop-in:
//     --passes N --checkpoint FILE [--resume]   progressive render, N primary-sample indices per pass (SailorPt_RenderProgressive)
//     --seed S --width W --device D --devices N     (N: CUDA devices to spread the frame over, SailorPtParams::deviceCount)
// Exit code: 0, or the negated SAILOR_PT_ERR_* code (the reference logs and returns, :94-98).
#include "../../include/sailor_pt.h"

#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

int main(int argc, char** argv)
{
	SailorPtParams p;
	memset(&p, 0, sizeof(p));
	p.height = 720; p.numSamples = 4; p.numAmbientSamples = 4; p.maxBounces = 4; p.msaa = 4;      // PathTracer.h:21-32 defaults; ambient (0,0,0)
	if (argc < 2)
	{
		fprintf(stderr, "usage: %s --in scene.gltf|.glb --out image.png|.pfm|.hdr [--height H] [--samples N] [--bounces B] [--camera NAME] [--ambient RRGGBB]\n"
			"       [--passes N --checkpoint FILE [--resume]] [--seed S] [--width W] [--device D] [--devices N]\n   backend: %s\n", argv[0], SailorPt_Backend());
		return 1;
	}
	int32_t rc = SailorPt_ParseCommandLineArgs(&p, const_cast<const char**>(argv), argc);
	if (rc != SAILOR_PT_OK) { fprintf(stderr, "bad arguments (%d): %s\n", rc, SailorPt_LastError()); return -rc; }
	p.numAmbientSamples = p.numSamples;                   // the reference never sets m_numAmbientSamples from the command line
	uint32_t passes = 0; const char* checkpoint = nullptr; bool resume = false; int32_t device = 0;
	for (int i = 1; i < argc; i++)
	{
		const std::string a = argv[i];
		if (a == "--passes" && i + 1 < argc) passes = (uint32_t)atoi(argv[++i]);
		else if (a == "--checkpoint" && i + 1 < argc) checkpoint = argv[++i];
		else if (a == "--resume") resume = true;
		else if (a == "--seed" && i + 1 < argc) p.seed = strtoull(argv[++i], nullptr, 10);
		else if (a == "--width" && i + 1 < argc) p.widthOverride = (uint32_t)atoi(argv[++i]);
		else if (a == "--device" && i + 1 < argc) device = atoi(argv[++i]);
		else if (a == "--devices" && i + 1 < argc) p.deviceCount = atoi(argv[++i]);
	}
	if (!p.pathToModel || !p.pathToModel[0]) { fprintf(stderr, "--in is required\n"); return 1; }
	rc = SailorPt_SetDevice(device);
	if (rc != SAILOR_PT_OK) { fprintf(stderr, "%s\n", SailorPt_LastError()); return -rc; }
	if (!passes && !checkpoint)
	{
		rc = SailorPt_Run(&p);                               // PathTracer::Run
	}
	else
	{
		SailorPtScene* scene = nullptr;
		rc = SailorPt_SceneLoad(p.pathToModel, &scene);
		if (rc == SAILOR_PT_OK)
		{
			uint32_t done = 0;
			rc = SailorPt_RenderProgressive(scene, &p, passes ? passes : p.msaa, 0, checkpoint, (resume ? 1u : 0u) | 2u, nullptr, nullptr, &done);
			if (rc == SAILOR_PT_OK) fprintf(stderr, "%u of %u primary-sample indices accumulated\n", done, p.msaa);
			SailorPt_SceneFree(scene);
		}
	}
	if (rc != SAILOR_PT_OK) { fprintf(stderr, "path tracer failed (%d): %s\n", rc, SailorPt_LastError()); return -rc; }
	SailorPtStats st;
	if (SailorPt_GetStats(&st) == SAILOR_PT_OK && st.secondsTotal > 0.0)
		fprintf(stderr, "%s: %.1f M rays in %.3f s (%.1f Mrays/s) on %u device(s)\n", SailorPt_Backend(), (double)st.rays / 1e6, st.secondsTotal, (double)st.rays / st.secondsTotal / 1e6, st.devicesUsed ? st.devicesUsed : 1u);
	return 0;
}
