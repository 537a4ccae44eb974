function compile_b200(model, varargin)
% COMPILE_B200  B200 replacement of @egdstmodel/compile.m, lines 754-819.
%
%   compile_b200(model)            % from inside compile.m, in place of the three `mex` calls
%   compile_b200(model, 'root', ROOT, 'python', 'python3')
%
% The reference's compile step writes modelspec.c/.h from the object's exec strings (compile.m:183-655) and
% builds three MEX files per model into model.dir (compile.m:754-819):
%       mex egdst_solver.c    egdst_lib.c modelspec.c   -> egdst_solver
%       mex egdst_call.c      egdst_lib.c modelspec.c   -> egdst_call
%       mex egdst_simulator.c egdst_lib.c modelspec.c   -> egdst_simulator
% This variant keeps everything before line 754 as it is (checks, optim_* inference, simlabels) and replaces
% the build:
%   1. the object's public properties are written as JSON (the generator's input);
%   2. `python -m egdst_b200.build_cli` restates the translation rules of compile.m:12-64,183-655 for CUDA
%      (egdst_b200/codegen.py) and runs nvcc for sm_100a: one shared library libegdst_b200_<key>.so per
%      generated model image, with the solver and simulator kernels specialised on the exec strings;
%   3. the three thin gateways of <root>/mex are built with `mex` against that library.  They have the
%      reference's names and arities (egdst_solver.c:149-154, egdst_simulator.c:54-75, egdst_call.c:45-58),
%      so egdstmodel.m (solve/sim/call), plot1-3.m and disp.m stay untouched.
% The -D flags (DISTRIB, TOLERANCE, ZEROCONSUMPTION, DOUBLEPOINT_DELTA, VERBOSE) are passed exactly as
% compile.m:757-777 builds them; the CUDA library reads the three tolerances at run time from the descriptor
% the gateways fill (mex/egdst_mex_common.h), DISTRIB is part of the generated image.
%
% Requirements on the MATLAB host: a B200 (sm_100a) with the CUDA toolkit (nvcc), Python >= 3.9 with numpy,
% and this repository at ROOT (default: two directories above this file).

p = inputParser;
p.addParameter('root', fileparts(fileparts(mfilename('fullpath'))));
p.addParameter('python', 'python3');
p.addParameter('force', false);
p.parse(varargin{:});
root = p.Results.root;
py = p.Results.python;

save_dir = pwd;
cleanup = onCleanup(@() cd(save_dir));
cd(model.dir);   % compile in the model directory, like compile.m:756

% ---- 1. the generator's input: the object's public properties --------------------------------------------
% jsonencode turns cell arrays into nested lists and struct arrays into lists of objects; EgdstModel.from_dict
% (egdst_b200/model.py) reads exactly that shape (tests/test_cpu_host.py holds a hand-written dump of
% model_retirement2.m in this form).
w = warning('off', 'MATLAB:structOnObject');
s = struct(model);
warning(w);
drop = {'M', 'D', 'sims', 'randstream', 'quadrature', 'lastrun_solver', 'lastrun_simulator', 'code'};
s = rmfield(s, intersect(fieldnames(s), drop));
jsonfile = fullfile(model.dir, 'model.json');
fid = fopen(jsonfile, 'w');
fwrite(fid, jsonencode(s));
fclose(fid);

% ---- 2. code generation + nvcc (replaces compile.m:183-655 and the model-specific half of :781,793,805) ---
cmd = sprintf('cd "%s" && "%s" -m egdst_b200.build_cli "%s" --out "%s"', root, py, jsonfile, model.dir);
if p.Results.force, cmd = [cmd ' --force']; end
if ~model.quiet, fprintf('Generating and compiling the CUDA model image'); tic; end
[st, out] = system(cmd);
if st ~= 0
    error('egdstmodel:compile', 'Error(s) while compiling the CUDA model image:\n%s', out);
end
lines = strsplit(strtrim(out), newline);
info = jsondecode(lines{end});       % {"key": ..., "library": ..., "optim": {...}}
if ~model.quiet, fprintf(' done in %s\n', model.ht(toc)); end

% the optim_* switches are compiled into the image; they must be the ones compile.m:669-747 inferred
fn = fieldnames(info.optim);
for i = 1:numel(fn)
    if logical(info.optim.(fn{i})) ~= logical(model.optim.(fn{i}))
        error('egdstmodel:compile', 'optim switch %s differs between compile.m and the CUDA generator', fn{i});
    end
end

% ---- 3. the three gateways (compile.m:757-819 with the reference sources replaced) ------------------------
switch model.shock.type
case 'lognormal'
    flags = ' -DDISTRIB=1';
case 'normal'
    flags = ' -DDISTRIB=2';
end
for tag = fieldnames(model.cflags)'
    val0 = model.cflags.(tag{1});
    if isnumeric(val0) && val0 == 0
        val1 = '0';
    elseif isnumeric(val0)
        val1 = sprintf('%1.0d', val0);
    else
        val1 = val0;
    end
    flags = [flags ' -D' tag{1} '=' val1]; %#ok<AGROW>
end
if ~model.quiet, fprintf('Compiler flags: %s\n', flags); end
[libdir, libname] = fileparts(info.library);
libname = regexprep(libname, '^lib', '');
gateways = {'egdst_solver', 'egdst_call', 'egdst_simulator'};
try
    for g = gateways
        runstr = ['mex ' fullfile(root, 'mex', [g{1} '.c']) ' -I' fullfile(root, 'include') ' -I' fullfile(root, 'mex') ...
                  flags ' -L' libdir ' -l' libname ' LDFLAGS=''$LDFLAGS -Wl,-rpath,' libdir ''' -outdir ' model.dir];
        if ~model.quiet, fprintf('Compiling %s.c', g{1}); tic; end
        eval(runstr);
        if ~model.quiet, fprintf(' done in %s\n', model.ht(toc)); end
    end
catch exception
    error('egdstmodel:compile', 'Error(s) while compiling C code!!!\n%s', exception.message);
end

% clear to require a new solve, mark the successful compile (compile.m:822-826)
model.M = {};
model.D = {};
model.sims = [];
model.needtocompile = false;
end
